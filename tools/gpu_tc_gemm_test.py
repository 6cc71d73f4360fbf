"""tcgen05 TF32 GEMM vs torch (run on the GPU box). Non-asserting table + structured error dump for debugging."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lsdm_b200.engine import Engine

torch.backends.cuda.matmul.allow_tf32 = False


def tf32_round(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def run(eng, M, N, K, act, group_max, bias_mode=1, exact=True, lda_pad=0, precision=1):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K + lda_pad, device="cuda", generator=g)[:, :K]
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(M if bias_mode == 2 else N, device="cuda", generator=g)
    if exact:
        A = tf32_round(A.contiguous()) if lda_pad == 0 else A.copy_(tf32_round(A.contiguous()))
        W = tf32_round(W)
    ref = A.double() @ W.double().T + (b.double()[:, None] if bias_mode == 2 else b.double())
    if act == 1:
        ref = ref.relu()
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    elif act == 3:
        ref = ref.sigmoid()
    if group_max:
        ref = ref.view(M // 32, 32, N).max(1)[0]
    got = eng.debug_gemm(A, W, b, act=act, group_max=group_max, bias_mode=bias_mode, precision=precision)
    torch.cuda.synchronize()
    err = (got.double() - ref).abs()
    rel = float((got.double() - ref).norm() / ref.norm())
    tag = f"M={M} N={N} K={K} act={act} gmax={int(group_max)} bm={bias_mode} exact={int(exact)} pad={lda_pad} prec={precision}"
    print(f"{tag:70s} rel_l2 {rel:.3e} max_abs {float(err.max()):.3e} nan {int(torch.isnan(got).sum())}")
    if rel > 1e-2 and not group_max:
        # structure of the error: which rows / cols are wrong
        bad = err > 1e-2 * float(ref.abs().max())
        rows = bad.any(1).nonzero().flatten()[:16].tolist()
        cols = bad.any(0).nonzero().flatten()[:16].tolist()
        print("   bad rows", rows, "bad cols", cols, "frac bad", float(bad.float().mean()))
        # try to identify a permutation: for row 0, find which ref column each got column matches
        r0 = got[0].double()
        match = [(int((ref[0] - v).abs().argmin()), float((ref[0] - v).abs().min())) for v in r0[:16]]
        print("   got[0,:16] best-matching ref cols:", match)
        c0 = got[:, 0].double()
        match = [(int((ref[:, 0] - v).abs().argmin()), float((ref[:, 0] - v).abs().min())) for v in c0[:16]]
        print("   got[:16,0] best-matching ref rows:", match)
    return rel


def main():
    eng = Engine(1)
    shapes = [(128, 32, 32), (256, 64, 64), (1000, 64, 64), (4096, 128, 128), (513, 192, 256), (300, 256, 512), (2048, 512, 256),
              (1024, 1024, 512), (100000, 64, 32), (128, 256, 32)]
    for (M, N, K) in shapes:
        for act in (0, 1):
            try:
                run(eng, M, N, K, act, False)
            except Exception:
                traceback.print_exc()
    for (M, N, K) in [(4096, 64, 32), (8192, 128, 64), (2048, 256, 128), (512, 512, 256), (32 * 37, 64, 32)]:
        try:
            run(eng, M, N, K, 1, True)
        except Exception:
            traceback.print_exc()
    run(eng, 1024, 256, 512, 2, False, bias_mode=2)
    run(eng, 4096, 128, 64, 3, False, lda_pad=192)
    run(eng, 4096, 128, 128, 2, False, exact=False)
    run(eng, 4096, 256, 256, 0, False, exact=False)
    for (M, N, K) in [(4096, 128, 128), (4096, 256, 256), (1000, 64, 32), (2048, 192, 256), (1024, 1024, 512), (4096, 32, 32)]:
        run(eng, M, N, K, 2, False, exact=False, precision=2)
    run(eng, 8192, 128, 64, 1, True, exact=False, precision=2)
    # timing: big memory-bound and compute-bound cases
    for (M, N, K, gm) in [(18874368 // 4, 32, 32, False), (18874368 // 4, 64, 32, True), (1 << 20, 128, 128, False), (1 << 19, 256, 256, False),
                          (65536, 1024, 512, False)]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
        for tf32 in (0, 1, 2):
            eng.debug_gemm(A, W, b, act=1, group_max=gm, precision=tf32)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                eng.debug_gemm(A, W, b, act=1, group_max=gm, precision=tf32)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            flops = 2.0 * M * N * K
            byts = 4.0 * (M * K + N * K + (M // 32 if gm else M) * N)
            print(f"time M={M} N={N} K={K} gmax={int(gm)} prec={int(tf32)}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
