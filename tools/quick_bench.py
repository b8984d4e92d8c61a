"""Quick forward/backward timing of ours vs the reference GPU build on one case (dev tool)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases, refgpu

def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    for view in (None, 2):
        c = cases.f3d_case(0, S, res, view, device="cuda")
        dL = cases.grad_seed(c)
        ours = refgpu.OursRun()
        o = ours.forward(c, decode_state=False)
        print(f"view={view} P={S*S} res={res} R={o['num_rendered']}")
        t_of = timeit(lambda: ours.forward(c, decode_state=False), iters)
        t_ob = timeit(lambda: ours.backward(c, dL), iters)
        print(f"  ours: fwd {t_of*1e3:.1f} us  bwd {t_ob*1e3:.1f} us")
        if refgpu.ref_available():
            ref = refgpu.RefRun()
            ref.forward(c, decode_state=False)
            t_rf = timeit(lambda: ref.forward(c, decode_state=False), iters)
            t_rb = timeit(lambda: ref.backward(c, dL), iters)
            print(f"  ref : fwd {t_rf*1e3:.1f} us  bwd {t_rb*1e3:.1f} us   speedup fwd {t_rf/t_of:.2f}x bwd {t_rb/t_ob:.2f}x")

if __name__ == "__main__":
    main()
