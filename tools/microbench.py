"""Micro-benchmark / profiling driver for single kernels at the live ASY-VRNet shapes.
    python tools/microbench.py [--case NAME ...] [--batch 8] [--iters 20] [--dtype bf16]
Prints one line per case: avg us (CUDA events, L2 flushed between launches), algorithmic GB/s and TFLOP/s.
Used under ncu (`ncu --set full -k regex:... python tools/microbench.py --case ... --iters 2`) for profiles/."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch  # noqa: E402
from vrcoc import ops  # noqa: E402
from vrcoc._lib import ACT_GELU, ACT_NONE  # noqa: E402

CASES = {
    # name: (kind, C, O, H, extra)
    "s1_fc1v": ("gnproj", 64, 256, 128, dict(split=128)),
    "s1_mlp1": ("gnproj", 64, 512, 128, dict(act=ACT_GELU)),
    "s1_mlp2": ("projres", 512, 64, 128, {}),
    "s1_fc2": ("projres", 128, 64, 128, {}),
    "s2_mlp1": ("gnproj", 128, 1024, 64, dict(act=ACT_GELU)),
    "s2_mlp2": ("projres", 1024, 128, 64, {}),
    "s3_fc1v": ("gnproj", 320, 512, 32, dict(split=256)),
    "s3_mlp1": ("gnproj", 320, 1280, 32, dict(act=ACT_GELU)),
    "s3_mlp2": ("projres", 1280, 320, 32, {}),
    "s4_mlp1": ("gnproj", 512, 2048, 16, dict(act=ACT_GELU)),
    "s4_mlp2": ("projres", 2048, 512, 16, {}),
    "s1_mlpf": ("mlpf", 64, 512, 128, {}),
    "s2_mlpf": ("mlpf", 128, 1024, 64, {}),
    "s3_mlpf": ("mlpf", 320, 1280, 32, {}),
    "s1_tmf": ("tmf", 64, 0, 128, {}),
    "s2_tmf": ("tmf", 128, 0, 64, {}),
    "s3_tmc": ("tmc", 320, 0, 32, {}),
    "s1_core": ("core", 128, 0, 128, dict(E=4, fold=8)),
    "s2_core": ("core", 128, 0, 64, dict(E=4, fold=4)),
    "s3_core": ("core", 256, 0, 32, dict(E=8, fold=2)),
    "s4_core": ("core", 256, 0, 16, dict(E=8, fold=1)),
    "n4_core": ("core", 96, 0, 32, dict(E=4, fold=2)),
    # the same launches through the run-time-geometry TMA kernel (VRCOC_CORE_FAST2=0), for A/B
    "s1_core_rt": ("core", 128, 0, 128, dict(E=4, fold=8, rt=True)),
    "s2_core_rt": ("core", 128, 0, 64, dict(E=4, fold=4, rt=True)),
    "s3_core_rt": ("core", 256, 0, 32, dict(E=8, fold=2, rt=True)),
    "s3_conv3": ("conv3", 320, 320, 32, {}),
    "s1_conv3": ("conv3", 64, 64, 128, {}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", nargs="*", default=list(CASES))
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--trace", action="store_true", help="tmf cases: print the per-CTA phase timeline (us) of one launch")
    args = ap.parse_args()
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    dev = "cuda"
    B = args.batch
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    es = 2 if dt == torch.bfloat16 else 4
    for name in args.case:
        kind, C, O, H, kw = CASES[name]
        P = H * H
        g = torch.Generator(device=dev).manual_seed(0)
        if kind == "gnproj":
            x = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            w = (torch.randn(O, C, device=dev, generator=g) / C ** 0.5).to(dt)
            bias = torch.zeros(O, device=dev)
            gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            _, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
            split = kw.get("split", 0)
            if split and dt == torch.bfloat16:
                out, out2 = torch.empty(B, split, H, H, device=dev), torch.empty(B, O - split, H, H, device=dev, dtype=dt)
                by = B * P * (C * es + split * 4 + (O - split) * es)
            else:
                out, out2 = torch.empty(B, O, H, H, device=dev, dtype=dt), None
                by = B * P * (C + O) * es
            d = ops.conv_desc(x, w, out, gn=(sums, gamma, beta, 1e-5), e_shift=bias, act=kw.get("act", ACT_NONE), out2=out2)
            fn = lambda: ops.conv_fwd(d)
            fl = 2.0 * B * P * C * O
        elif kind == "projres":
            x = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            w = (torch.randn(O, C, device=dev, generator=g) / C ** 0.5).to(dt)
            bias, ls = torch.zeros(O, device=dev), torch.ones(O, device=dev)
            res = torch.randn(B, O, H, H, device=dev, generator=g).to(dt)
            out = torch.empty_like(res)
            sums = ops.new_sample_sums(B, dev)
            d = ops.conv_desc(x, w, out, e_shift=bias, post_scale=ls, res=res, out_sample_sums=sums)
            fn = lambda: ops.conv_fwd(d)
            by, fl = B * P * (C + 2 * O) * es, 2.0 * B * P * C * O
        elif kind == "mlpf":
            x = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            w1 = (torch.randn(O, C, device=dev, generator=g) / C ** 0.5).to(dt)
            w2 = (torch.randn(C, O, device=dev, generator=g) / O ** 0.5).to(dt)
            b1, b2, ls = torch.zeros(O, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)
            gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            sums = ops.sample_sums_of(x)
            osum = ops.new_sample_sums(B, dev)
            fn = lambda: ops.mlp_fused_fwd(x, sums, gamma, beta, 1e-5, w1, b1, w2, b2, ls, osum)
            by, fl = B * P * 3 * C * es, 4.0 * B * P * C * O
        elif kind == "tmf":
            # fused token-mixer half (csrc/token_mixer_fused.cu): algorithmic bytes = x in + out out (SURVEY 8d, 2*C*P*s)
            x = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            ED = 128
            w1 = (torch.randn(ED, C, device=dev, generator=g) / C ** 0.5).to(dt)
            wv = (torch.randn(ED, C, device=dev, generator=g) / C ** 0.5).to(dt)
            w2 = (torch.randn(C, ED, device=dev, generator=g) / ED ** 0.5).to(dt)
            zED, zC = torch.zeros(ED, device=dev), torch.zeros(C, device=dev)
            w_fold, k0, k1 = ops.fold_gn_weights(w1, zED, wv, zED, torch.ones(C, device=dev), zC)
            sums = ops.sample_sums_of(x)
            osum = ops.new_sample_sums(B, dev)
            a, b_ = torch.ones(1, device=dev), torch.zeros(1, device=dev)
            fold = H // 16
            fn = lambda: ops.token_mixer_fused_fwd(x, sums, 1e-5, w_fold, k0, k1, a, b_, w2, zC, torch.ones(C, device=dev), osum, 4, 32, fold, fold)
            by, fl = B * P * 2 * C * es, B * P * (2.0 * C * ED * 4 + 13.0 * ED)
            nb_units = B * fold * fold
        elif kind == "tmc":
            # stage-3 projection + core (csrc/token_mixer_fused.cu, second kernel): x in, o out
            heads, ED = 8, 256
            x = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            w1 = (torch.randn(ED, C, device=dev, generator=g) / C ** 0.5).to(dt)
            wv = (torch.randn(ED, C, device=dev, generator=g) / C ** 0.5).to(dt)
            zED = torch.zeros(ED, device=dev)
            w_fold, k0, k1 = ops.fold_gn_weights(w1, zED, wv, zED, torch.ones(C, device=dev), torch.zeros(C, device=dev))
            sums = ops.sample_sums_of(x)
            a, b_ = torch.ones(1, device=dev), torch.zeros(1, device=dev)
            fn = lambda: ops.token_mixer_core_fwd(x, sums, 1e-5, w_fold, k0, k1, a, b_, heads, 32, 2, 2)
            nb_units = B * 4 * 2
            by, fl = B * P * (C + ED) * es, B * P * (2.0 * C * ED * 3 + 13.0 * ED)
        elif kind == "conv3":
            x = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            w = (torch.randn(O, C * 9, device=dev, generator=g) / (C * 9) ** 0.5).to(dt)
            out = torch.empty(B, O, H, H, device=dev, dtype=dt)
            d = ops.conv_desc(x, w, out, kh=3, kw=3, stride=1, pad=1)
            fn = lambda: ops.conv_fwd(d)
            by, fl = B * P * (C + O) * es, 2.0 * B * P * C * O * 9
        else:
            E, fold = kw["E"], kw["fold"]
            feat = torch.randn(B, C, H, H, device=dev, generator=g)
            value = torch.randn(B, C, H, H, device=dev, generator=g).to(dt)
            a, b_ = torch.ones(1, device=dev), torch.zeros(1, device=dev)
            if kw.get("rt"):
                def fn():
                    os.environ["VRCOC_CORE_FAST2"] = "0"
                    try:
                        ops.cluster_core_fwd(feat, value, a, b_, E, fold, fold, 2, 2)
                    finally:
                        del os.environ["VRCOC_CORE_FAST2"]
            else:
                fn = lambda: ops.cluster_core_fwd(feat, value, a, b_, E, fold, fold, 2, 2)
            by, fl = B * P * C * (4 + 2 * es), 13.0 * B * P * C
        if kind in ("tmf", "tmc") and args.trace:
            from vrcoc._lib import lib
            nb = min(148, nb_units)
            buf = torch.zeros(nb * 4 * 16, dtype=torch.int64, device=dev)
            lib.vrcoc_debug_set_tm_trace(buf.data_ptr())
            fn(); torch.cuda.synchronize()
            lib.vrcoc_debug_set_tm_trace(None)
            t = buf.cpu().reshape(nb, 4, 16).double()
            names = {0: "x+W1 landed", 1: "GEMM1 issued", 2: "GEMM1 done", 8: "w0: GEMM1 seen", 9: "w0: pass1", 10: "w0: pass2", 11: "w0: V conv",
                     7: "w0: agg seen", 12: "w0: pass4", 3: "o ready seen", 4: "GEMM2 done", 13: "w0: GEMM2 seen", 14: "w0: epilogue", 5: "epi seen",
                     6: "store issued", 15: "entry"}
            for cta in (0, nb // 2, nb - 1):
                t0 = t[cta, 0, 0]
                for it in range(4):
                    if t[cta, it, 0] == 0:
                        continue
                    order = sorted((k for k in names if t[cta, it, k] > 0), key=lambda k: t[cta, it, k])
                    print(f"  cta {cta} it {it}: " + "  ".join(f"{names[k]} {((t[cta, it, k] - t0) / 1e3).item():.2f}" for k in order))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.iters):
            if not args.no_flush:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        us = tot / args.iters * 1e3
        print(f"{name:10s} {us:9.1f} us   {by / us / 1e3:8.1f} GB/s   {fl / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
