"""On-device cross-check of the tcgen05 gather-GEMM against the fp32 SIMT kernel (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpd_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True
for (m_in, m_out, cin, cout, K) in [(300, 128, 32, 64, 27), (5000, 5000, 64, 64, 27), (70000, 70000, 128, 128, 27),
                                    (9000, 4000, 16, 32, 27), (141376, 141376, 256, 128, 9), (35344, 35344, 256, 256, 9),
                                    (1000, 777, 8, 16, 3), (141376, 141376, 512, 64, 9)]:
    x = torch.randn(m_in, cin, device=dev)
    w = torch.randn(cout, K, cin, device=dev) / (cin * K) ** 0.5
    nbr = torch.randint(-m_in, m_in, (m_out, K), device=dev, dtype=torch.int32).clamp_(min=-1)
    if m_out > 1000:
        nbr[128:256] = -1                 # a tile with no work at all
        nbr[:, 5] = -1                    # a tap nobody uses
    bias = torch.randn(cout, device=dev)
    scale, shift = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
    res = torch.randn(m_out, cout, device=dev)
    st_a, st_b = torch.zeros(2, cout, device=dev), torch.zeros(2, cout, device=dev)
    ref = ops.gather_gemm(x, w, nbr, bias=bias, scale=scale, shift=shift, residual=res, relu=True, stats=st_a, algo=ops.ALGO_SIMT)
    torch.cuda.synchronize()
    t0 = time.time()
    got = ops.gather_gemm(x, w, nbr, bias=bias, scale=scale, shift=shift, residual=res, relu=True, stats=st_b, algo=ops.ALGO_TCGEN05)
    torch.cuda.synchronize()
    x64, w64 = x.double(), w.double()
    ex = torch.zeros(m_out, cout, dtype=torch.float64, device=dev)
    for k in range(K):
        idx = nbr[:, k].long()
        ex += torch.where((idx >= 0)[:, None], x64[idx.clamp(min=0)], torch.zeros((), dtype=torch.float64, device=dev)) @ w64[:, k, :].T
    e_tc = (ops.gather_gemm(x, w, nbr, algo=ops.ALGO_TCGEN05).double() - ex).abs().max().item()
    e_simt = (ops.gather_gemm(x, w, nbr, algo=ops.ALGO_SIMT).double() - ex).abs().max().item()
    print(f"   vs fp64: tc {e_tc:.2e}  simt {e_simt:.2e}  (|y|max {ex.abs().max().item():.2f})", flush=True)
    err = (got - ref).abs().max().item()
    serr = ((st_a - st_b).abs() / (st_a.abs() + 1)).max().item()
    plain = (ops.gather_gemm(x, w, nbr, algo=ops.ALGO_TCGEN05) - ops.gather_gemm(x, w, nbr, algo=ops.ALGO_SIMT)).abs().max().item()
    # timing
    for algo, name in ((ops.ALGO_SIMT, "simt"), (ops.ALGO_TCGEN05, "tc")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gather_gemm(x, w, nbr, bias=bias, algo=algo)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        P = int((nbr >= 0).sum())
        print(f"   {name}: {ms*1e3:8.1f} us  {2.0*P*cin*cout/ms/1e9:7.1f} TFLOP/s (useful)", flush=True)
    good = e_tc < 1e-4 and err < 2e-4 and serr < 1e-3
    ok &= good
    print(f"m_out={m_out} cin={cin} cout={cout} K={K}: max|tc-simt| fused={err:.2e} plain={plain:.2e} stats={serr:.2e} {'OK' if good else 'FAIL'}", flush=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
