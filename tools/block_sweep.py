"""BASELINE.json configs[1]: single context-cluster block micro-benchmark sweep, forward and forward+backward, on one B200.
    python tools/block_sweep.py [--batch 1|8|16|64] [--iters 20] [--dtype bf16|fp32] [--quick] [--out profiles/xxx.txt]
    python bench.py --mode block-sweep [-- same flags]        (the same sweep, reachable from the bench entry point)
One line per configuration: the seven live rows of SURVEY 8 (S1..N3) and a grid C x HxW x centres x (E,D) around them.
Columns: algorithmic bytes of a whole ClusterBlock = 4*C*P*s forward (two round trips: GroupNorm 2 needs the
finished first half), 6*C*P*s backward; flops = 6*C*ED*P + (2M+5)*ED*P + 4*r*C^2*P forward, 2x that backward.
`hbm%` = bytes / time / MEASURED hbm_gbs, `tc%` = flops / time / sustained bf16 TFLOP/s; the block is bound by the larger.
Timing: CUDA events around one block call (all its launches) queued behind a stream hold so that host launch overhead is
not measured, L2 flushed before every call, after 3 warm-up calls.  `fwd` is the gradient-free path (torch.no_grad),
`fwd+bwd` the autograd path (native backward kernels)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch  # noqa: E402
import vrcoc  # noqa: E402

LIVE = [  # name, C, H, fold, proposal, heads, head_dim, mlp_ratio   (live configurations at phi='l', 512x512)
    ("S1", 64, 128, 8, 2, 4, 32, 8), ("S2", 128, 64, 4, 2, 4, 32, 8), ("S3", 320, 32, 2, 2, 8, 32, 4), ("S4", 512, 16, 1, 2, 8, 32, 4),
    ("N5", 512, 16, 2, 2, 4, 24, 4), ("N4", 640, 32, 2, 2, 4, 24, 4), ("N3", 256, 64, 2, 2, 4, 24, 4),
]


def peaks():
    p = {"hbm_gbs": 6461.2, "bf16_tflops": 1407.5}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p["hbm_gbs"] = m.get("hbm_gbs", p["hbm_gbs"])
        p["bf16_tflops"] = m.get("bf16_tflops_sustained", p["bf16_tflops"])
    except OSError:
        pass
    return p


def flops_bytes(C, H, heads, hd, ratio, prop, B, es):
    """SURVEY 8d per-unit figures of a whole ClusterBlock forward, times B."""
    P, ED, M = H * H, heads * hd, prop * prop
    by_f = 4.0 * C * P * es * B
    fl_f = (6.0 * C * ED * P + (2 * M + 5) * ED * P + 4.0 * ratio * C * C * P) * B
    return by_f, fl_f


class Timer:
    def __init__(self, iters, dev="cuda"):
        self.iters = iters
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def __call__(self, fn, hold_ms):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(self.iters):
            self.flush.zero_()
            # the stream is held busy while the host queues the block's launches (eager Python costs more per launch than
            # the small kernels run): the events then bracket back-to-back GPU work, not host launch gaps
            torch.cuda._sleep(int(1.9e6 * hold_ms))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        return tot / self.iters * 1e3


def time_case(case, B, dt, timed, backward=True, dev="cuda"):
    """-> dict(us_fwd, us_fwd_bwd, bytes_fwd, flops_fwd) for one (name, C, H, fold, prop, heads, hd, ratio) case"""
    name, C, H, fold, prop, heads, hd, ratio = case
    es = 2 if dt == torch.bfloat16 else 4
    torch.manual_seed(0)
    blk = vrcoc.ClusterBlock(dim=C, mlp_ratio=float(ratio), proposal_w=prop, proposal_h=prop, fold_w=fold, fold_h=fold, heads=heads,
                             head_dim=hd).to(dev, dt)
    with torch.no_grad():
        blk.layer_scale_1.uniform_(0.5, 1.5)
        blk.layer_scale_2.uniform_(0.5, 1.5)
    x = torch.randn(B, C, H, H, device=dev).to(dt)
    g = torch.randn(B, C, H, H, device=dev).to(dt)
    by_f, fl_f = flops_bytes(C, H, heads, hd, ratio, prop, B, es)

    def fwd():
        with torch.no_grad():
            blk(x)

    xg = x.clone().requires_grad_(True)

    def fwd_bwd():
        for p in blk.parameters():
            p.grad = None
        xg.grad = None
        blk(xg).backward(g)

    out = {"us_fwd": timed(fwd, 1.0), "bytes_fwd": by_f, "flops_fwd": fl_f}
    if backward:
        out["us_fwd_bwd"] = timed(fwd_bwd, 6.0)
    return out


def live_rows(B, dt, iters=10, backward=False):
    """the seven live rows at batch B: {name: {us_fwd, hbm_frac, tc_frac, ...}} (bench.py's `roofline_block`)"""
    pk = peaks()
    timed = Timer(iters)
    res = {}
    for case in LIVE:
        r = time_case(case, B, dt, timed, backward=backward)
        t = r["us_fwd"]
        res[case[0]] = {"us_fwd": round(t, 2), "alg_MB": round(r["bytes_fwd"] / 1e6, 2), "alg_GFLOP": round(r["flops_fwd"] / 1e9, 3),
                        "hbm_frac": round(r["bytes_fwd"] / t / 1e3 / pk["hbm_gbs"], 4),
                        "tc_frac": round(r["flops_fwd"] / t / 1e6 / pk["bf16_tflops"], 4)}
        if backward:
            res[case[0]]["us_fwd_bwd"] = round(r["us_fwd_bwd"], 2)
    return res


def sweep_cases(quick, B, es):
    cases = list(LIVE)
    if not quick:
        for (E, D) in ((4, 32), (8, 32), (4, 24)):
            for C in (64, 128, 256, 320, 512, 640):
                for H in (16, 32, 64, 128):
                    for prop in (2, 3, 4):
                        if C * H * H * B * es > 600e6:           # keep the largest activations well under a GB
                            continue
                        if (E, D) != (4, 32) and (prop != 2 or C not in (128, 512)):   # the (E,D) axis on a sub-grid
                            continue
                        tag = "" if (E, D) == (4, 32) else f"e{E}d{D}"
                        cases.append((f"g{C}x{H}m{prop * prop}{tag}", C, H, H // 16, prop, E, D, 4))
    return cases


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--quick", action="store_true", help="live rows only")
    ap.add_argument("--no-backward", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    es = 2 if dt == torch.bfloat16 else 4
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pk = peaks()
    B = args.batch
    cases = sweep_cases(args.quick, B, es)
    timed = Timer(args.iters)
    lines = [f"# ClusterBlock sweep, batch {B}, {args.dtype}; peaks: {pk['hbm_gbs']:.0f} GB/s, {pk['bf16_tflops']:.0f} TFLOP/s (bf16, sustained)",
             f"# {'case':16s} {'C':>4s} {'HxW':>8s} {'reg':>4s} {'M':>3s} {'ExD':>5s} | {'fwd us':>8s} {'hbm%':>6s} {'tc%':>6s} | {'fwd+bwd us':>10s} {'hbm%':>6s} {'tc%':>6s}"]
    print("\n".join(lines), flush=True)
    for case in cases:
        name, C, H, fold, prop, heads, hd, ratio = case
        r = time_case(case, B, dt, timed, backward=not args.no_backward)
        by_f, fl_f, t_f = r["bytes_fwd"], r["flops_fwd"], r["us_fwd"]
        t_fb = r.get("us_fwd_bwd", float("nan"))
        by_fb, fl_fb = by_f + 6.0 * C * H * H * es * B, 3.0 * fl_f
        row = (f"  {name:16s} {C:4d} {H:4d}x{H:<3d} {H // max(fold, 1):4d} {prop * prop:3d} {heads}x{hd:<3d} | {t_f:8.1f} {100 * by_f / t_f / 1e3 / pk['hbm_gbs']:6.1f} "
               f"{100 * fl_f / t_f / 1e6 / pk['bf16_tflops']:6.1f} | {t_fb:10.1f} {100 * by_fb / t_fb / 1e3 / pk['hbm_gbs']:6.1f} "
               f"{100 * fl_fb / t_fb / 1e6 / pk['bf16_tflops']:6.1f}")
        print(row, flush=True)
        lines.append(row)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
