import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
base = dict(gpus=1, steps=2, warmup=3, impl="b200", workload="c5", long_mesh="auto", dtype="f64", no_cpu=True, no_extra=True)
for name in sys.argv[1:]:
    a = argparse.Namespace(**{**base, "workload": name})
    if name in bench.SINGLE:
        r = bench.run_single_gpu_arm(a, emit=False, sampler=False)
        print(name, "ms", round(r["ms_per_step"], 2), "kernel_ms", round(r["roofline"]["kernel_ms"], 2), "e2e %.3e" % r["e2e"]["value"], flush=True)
    else:
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            bench.run_gpu_arm(a)
        print(name, "done", flush=True)
    print("mem allocated GB", torch.cuda.memory_allocated() / 1e9, "reserved", torch.cuda.memory_reserved() / 1e9, flush=True)
