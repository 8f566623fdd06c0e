"""Collects the reference's own golden vectors / fixtures for the pose-optimisation path into tests/golden/.

Run in the build container (where /root/reference is mounted); the GPU box only sees the committed copies.
Sources (DLR-RM/3DObjectTracking @ f0210618, M3T/):
  data/modality_test/{region_modality_global,region_modality_local,depth_modality}_{gradient,hessian}.txt
      known answers of RegionModalityTest.Calculate{Global,Local}GradientAndHessian, DepthModalityTest.
      CalculateGradientAndHessian (test/modality_test.cpp:280-316,534-550), tolerance 1e-3 relative
  data/{optimizer,tracker,refiner}_test/triangle_pose.txt   pose known answers (1e-5 relative)
  data/model_test/{region,depth}_model.bin                  checked-in sparse viewpoint models (162 views x 10 points)
  data/_sequence/{color,depth}_camera_image_200.png + camera yaml values, data/_body/triangle.obj
      the real frame pair and body the known answers were computed on
These are data fixtures (numbers, images), not source code.
"""
import json
import os
import shutil

REF = "/root/reference/M3T"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_matrix_txt(path):
    lines = open(path).read().strip().split("\n")
    rows, cols = [int(x) for x in lines[1].replace("\t", "").split(",") if x.strip()]
    vals = []
    for ln in lines[2:2 + rows]:
        vals.append([float(x) for x in ln.replace("\t", "").split(",") if x.strip()])
    return {"name": lines[0].strip(), "rows": rows, "cols": cols, "data": vals}


def main():
    out = {}
    for stem in ("region_modality_global", "region_modality_local", "depth_modality"):
        for kind in ("gradient", "hessian"):
            out[f"{stem}_{kind}"] = read_matrix_txt(f"{REF}/data/modality_test/{stem}_{kind}.txt")
    for t in ("optimizer", "tracker", "refiner"):
        out[f"{t}_triangle_pose"] = read_matrix_txt(f"{REF}/data/{t}_test/triangle_pose.txt")
    out["detector_triangle_pose"] = read_matrix_txt(f"{REF}/data/detector_test/detector_triangle_pose.txt")
    # fixtures of the test rig (test/common_test.cpp:9-13, data/_sequence/*.yaml, data/optimizer_test/optimizer.yaml)
    out["triangle_world2body"] = [[0.607676, 0.408914, -0.680823, 0.472944], [0.786584, -0.428213, 0.444880, -0.213009],
                                  [-0.109620, -0.805867, -0.581860, 0.346384], [0, 0, 0, 1]]
    out["color_camera"] = {"fu": 698.128, "fv": 698.617, "ppu": 478.459, "ppv": 274.426, "width": 960, "height": 540,
                           "camera2world": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]}
    out["depth_camera"] = {"fu": 425.773, "fv": 425.773, "ppu": 427.202, "ppv": 237.662, "width": 848, "height": 480,
                           "depth_scale": 0.001,
                           "camera2world": [[0.99985489, 0.00778240, 0.01509715, 0.01453388],
                                            [-0.00782678, 0.99996543, 0.00288261, 0.00013995],
                                            [-0.01507424, -0.00300036, 0.99988175, 0.00051057], [0, 0, 0, 1]]}
    out["optimizer_test_tikhonov"] = {"rotation": 5000.0, "translation": 500000.0}
    out["triangle_obj"] = {"vertices": [[-0.038305, 0.0, 0.0], [-0.038305, 0.0, 0.012], [0.019152, -0.033231, 0.0],
                                        [0.019152, -0.033231, 0.012], [0.019152, 0.033231, 0.0], [0.019152, 0.033231, 0.012]],
                           "faces": [[1, 3, 4], [3, 5, 4], [4, 6, 2], [5, 1, 2], [1, 5, 3], [2, 1, 4], [5, 6, 4], [6, 5, 2]],
                           "geometry2body_translation": [0.0, 0.0, -0.006]}
    json.dump(out, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1)
    for f in ("region_model.bin", "depth_model.bin"):
        shutil.copyfile(f"{REF}/data/model_test/{f}", os.path.join(HERE, f))
    for f in ("color_camera_image_200.png", "depth_camera_image_200.png"):
        shutil.copyfile(f"{REF}/data/_sequence/{f}", os.path.join(HERE, f))
    os.chmod(os.path.join(HERE, "region_model.bin"), 0o644)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
