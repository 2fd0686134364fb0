"""On-device cross-check of the tcgen05 weight-gradient against the SIMT kernel and fp64 (run under `timeout`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpd_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True
cases = [(3000, 2500, 32, 32, 27, 0.3), (300, 200, 16, 16, 27, 0.5), (60000, 60000, 64, 64, 27, 0.4), (70000, 70000, 128, 128, 27, 0.5),
         (141376, 141376, 256, 128, 9, 0.97), (35344, 35344, 256, 256, 9, 0.97), (141376, 141376, 512, 64, 9, 0.97),
         (35344, 141376, 128, 256, 1, 1.0), (5000, 9000, 16, 32, 27, 0.2), (141376, 141376, 64, 64, 9, 0.97)]
for (m_in, m_out, cin, cout, K, dens) in cases:
    x = torch.randn(m_in, cin, device=dev)
    dy = torch.randn(m_out, cout, device=dev)
    nbr = torch.randint(0, m_in, (m_out, K), device=dev, dtype=torch.int32)
    nbr[torch.rand(m_out, K, device=dev) > dens] = -1
    if K > 5:
        nbr[:, 5] = -1
    ref, _ = ops.gather_wgrad(x, dy, nbr, algo=ops.ALGO_SIMT)
    nbr_t = nbr.t().contiguous()
    got, _ = ops.gather_wgrad(x, dy, nbr.t().contiguous(), algo=ops.ALGO_TCGEN05, tap_major=True)
    nbr_t = nbr.t().contiguous()
    got_t, _ = ops.gather_wgrad(x, dy, nbr_t, algo=ops.ALGO_TCGEN05, tap_major=True)
    ref_t, _ = ops.gather_wgrad(x, dy, nbr_t, algo=ops.ALGO_SIMT, tap_major=True)
    torch.cuda.synchronize()
    ex = torch.zeros(cout, K, cin, dtype=torch.float64, device=dev)
    for k in range(K):
        idx = nbr[:, k].long()
        xg = torch.where((idx >= 0)[:, None], x.double()[idx.clamp(min=0)], torch.zeros((), dtype=torch.float64, device=dev))
        ex[:, k, :] = dy.double().T @ xg
    scale = max(1.0, ex.abs().max().item())
    e_tc, e_simt = (got.double() - ex).abs().max().item() / scale, (ref.double() - ex).abs().max().item() / scale
    for algo, name in ((ops.ALGO_SIMT, "simt"), (ops.ALGO_TCGEN05, "tc")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.gather_wgrad(x, dy, nbr_t, algo=algo, tap_major=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        P = int((nbr >= 0).sum())
        print(f"   {name}: {ms*1e3:9.1f} us  {2.0*P*cin*cout/ms/1e9:7.1f} TFLOP/s (useful)  {P*(cin+cout)*4/ms/1e6:7.0f} GB/s gathered", flush=True)
    e_t = max((got_t.double() - ex).abs().max().item(), (ref_t.double() - ex).abs().max().item()) / scale
    for name, tbl, tm in (("tc tap-major", nbr_t, True),):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.gather_wgrad(x, dy, tbl, algo=ops.ALGO_TCGEN05, tap_major=tm)
        e1.record(); torch.cuda.synchronize()
        print(f"   {name}: {e0.elapsed_time(e1) / 3 * 1e3:9.1f} us", flush=True)
    good = e_tc < 1e-4 and e_t < 1e-4
    ok &= good
    print(f"m_out={m_out} cin={cin} cout={cout} K={K}: rel err vs fp64 tc={e_tc:.2e} simt={e_simt:.2e} (|dw|max {scale:.1f}) {'OK' if good else 'FAIL'}", flush=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
