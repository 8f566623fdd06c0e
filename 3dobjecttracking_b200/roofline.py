"""Algorithmic-byte accounting of the fused kernel (SURVEY.md §8d / BASELINE.md §4).

Compulsory, cache-perfect traffic per body per correspondence iteration at scale s:
    region: n_lines * (32 + 57*s) + 8*n_bins^3 + 12*n_views + 96
    depth : n_points * (24 + 2*(ns+1)^2) + 12*n_views          (ns = realised n_strides)
LUT and view-orientation terms are charged once per body per iteration although the fused kernel keeps them
on chip / in L2 across iterations, so a fraction above 1.0 of this accounting is possible.
"""
from __future__ import annotations

import numpy as np


def _last_valid(v, i):
    return v[i] if i < len(v) else v[-1]


def depth_n_strides(fu, z, considered_distance, stride_length):
    """Realised n_strides of DepthModality::FindCorrespondence (depth_modality.cpp:832-836) at depth z."""
    max_n = int(considered_distance / stride_length + 0.5)
    diameter = 2.0 * considered_distance * fu / z
    stride = int(diameter / max_n + 1.0)
    return int(diameter / stride + 0.5)


def algorithmic_bytes_per_step(wl, n_corr=None, body_depths=None):
    """Bytes for one tracking step (n_corr correspondence iterations) of ALL bodies of the workload.
    Returns (total_bytes, region_bytes, depth_bytes, line_evals, point_evals)."""
    n_corr = wl.n_corr_iterations if n_corr is None else n_corr
    nb = wl.n_bodies
    region_b = depth_b = 0.0
    line_evals = point_evals = 0
    if wl.region:
        nv = wl.region_model.n_views
        for c in range(n_corr):
            s = _last_valid(wl.region.scales, c)
            region_b += nb * (wl.region.n_lines_max * (32 + 57 * s) + 8 * wl.region.n_histogram_bins ** 3 + 12 * nv + 96)
            line_evals += nb * wl.region.n_lines_max
    if wl.depth:
        nv = wl.depth_model.n_views
        if body_depths is None:
            w2c = np.asarray(wl.depth_world2camera, np.float64)
            body_depths = [float(w2c[2, :3] @ wl.gt_body2world[b][:, 3] + w2c[2, 3]) for b in range(nb)]
        for c in range(n_corr):
            d = _last_valid(wl.depth.considered_distances, c)
            for z in body_depths:
                ns = depth_n_strides(wl.depth_intrinsics.fu, z, d, wl.depth.stride_length)
                depth_b += wl.depth.n_points_max * (24 + 2 * (ns + 1) ** 2) + 12 * nv
            point_evals += nb * wl.depth.n_points_max
    return region_b + depth_b, region_b, depth_b, line_evals, point_evals
