import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, "oracle"))
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat"):
    try: print(f, open(f).read().replace("\n", " | ")[:300])
    except Exception as e: print(f, e)
print({k: v for k, v in os.environ.items() if k.startswith("OMP") or k.startswith("GOMP") or k.startswith("KMP")})
if len(sys.argv) > 1 and sys.argv[1] == "torch":
    import torch
    print("torch threads", torch.get_num_threads())
import oracle_py
pkg = importlib.import_module("3dobjecttracking_b200")
wl = pkg.synth.make_workload("c4", n_bodies=128, n_divides=4)
for native in (True,):
    for n in (1, 8, 16, 32, 64, 128):
        trk = oracle_py.OracleTracker(wl, n_threads=n, native=native)
        trk.tracking_step(0)
        ts = []
        for _ in range(3):
            trk.set_poses(wl.start_body2world)
            t0 = time.perf_counter(); trk.tracking_step(0); ts.append(time.perf_counter() - t0)
        print("threads", n, "ms/step", [round(1e3 * t, 1) for t in ts], "it/s", round(128 * 7 / min(ts)))
