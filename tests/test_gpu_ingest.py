"""Frame-ingest lifetime rules: a pinned frame stays referenced by its camera until the next upload or
m3tb_detach_frames; after detaching, the caller may overwrite the buffer and every later call still sees the frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_detach_frames_gives_the_pinned_buffers_back(capi, synth):
    import torch
    wl = synth.make_workload("c2", n_bodies=4, n_divides=3, seed=23)
    pin_c = torch.from_numpy(wl.color_frames.copy()).pin_memory()
    pin_d = torch.from_numpy(wl.depth_frames.view(np.uint8).reshape(wl.n_bodies, wl.depth_frames.shape[1], -1).copy()).pin_memory()
    out = []
    for pinned in (False, True):
        ctx = capi.context_from_workload(wl, upload_frames=not pinned)
        if pinned:
            ctx.upload_batch_ptr(True, 0, wl.n_bodies, pin_c.data_ptr(), pin_c.stride(0), pin_c.stride(1))
            ctx.upload_batch_ptr(False, 0, wl.n_bodies, pin_d.data_ptr(), pin_d.stride(0), pin_d.stride(1))
        ctx.start_modalities(0)
        ctx.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations)
        if pinned:
            assert 0 < ctx.last_ingest_bytes() < wl.color_frames.nbytes + wl.depth_frames.nbytes  # ROI-only so far
            ctx.detach_frames()
            pin_c.fill_(7)      # the caller reuses its buffers
            pin_d.fill_(9)
        ctx.calculate_results(0)                       # histogram update walks the frame again
        ctx.tracking_step(1, wl.n_corr_iterations, wl.n_update_iterations)   # and so does a second step
        hf, hb = ctx.get_histograms(0, wl.region.n_histogram_bins)
        out.append((ctx.get_poses(), hf, hb))
        ctx.close()
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
