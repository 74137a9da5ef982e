"""GPU timeline of one replay of the benched inference graph (phi='l', 512x512, batch 8, bf16): how much of the step the GPU is
busy, how many kernels overlap, which kernels sit on the critical path and where the gaps are.
    python tools/timeline.py [--batch 8] [--out profiles/rNN_timeline.txt]
CUPTI kernel records through torch.profiler; a number printed here is a profile, never a bench value."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch  # noqa: E402
import vrcoc  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def short(n):
    n = n.replace("void ", "").replace("vrcoc::", "").replace("(anonymous namespace)::", "")
    return n.split("(")[0][:56]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--phi", default="l")
    ap.add_argument("--out", default=None)
    ap.add_argument("--list", action="store_true", help="append the ordered kernel list (start, duration, name)")
    args = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    model = vrcoc.EfficientVRNet(num_classes=4, num_seg_classes=9, phi=args.phi).eval().to(dev, torch.bfloat16)
    sess = vrcoc.InferenceSession(model, batch=args.batch, img=512, slots=1, decode=True)
    S = sess.slots[0]
    for _ in range(3):
        S["graph"].replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        S["graph"].replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ks = sorted(((e.time_range.start, e.time_range.end, short(e.name)) for e in ev), key=lambda t: t[0])
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    lines = []
    P = lines.append
    P(f"# one graph replay, batch {args.batch}: {len(ks)} kernels, span {(t1 - t0):.1f} us, sum of kernel durations {sum(k[1] - k[0] for k in ks):.1f} us")
    # union busy time and concurrency histogram
    pts = sorted([(k[0], 1) for k in ks] + [(k[1], -1) for k in ks])
    depth, last, busy = 0, t0, collections.Counter()
    for t, d in pts:
        busy[depth] += t - last
        last, depth = t, depth + d
    P("# time with n kernels in flight: " + "  ".join(f"n={n}: {v:.0f} us ({100 * v / (t1 - t0):.1f}%)" for n, v in sorted(busy.items())))
    # gaps (nothing running)
    gaps, cur_end, prev = [], ks[0][1], ks[0][2]
    for s, e, n in ks[1:]:
        if s > cur_end:
            gaps.append((s - cur_end, prev, n, cur_end - t0))
        if e > cur_end:
            cur_end, prev = e, n
    P(f"# idle gaps: {len(gaps)}, total {sum(g[0] for g in gaps):.1f} us; the 12 longest:")
    P("# (the ~190 us hole after the first ~10 kernels is the tail of the graph LAUNCH of this isolated replay - the head of the graph runs "
      "while the rest is still being submitted; tools/gap_probe.py replays the same kernels captured alone without it, and back-to-back "
      "replays hide it behind the previous step)")
    for g in sorted(gaps, reverse=True)[:12]:
        P(f"    {g[0]:6.1f} us at t={g[3]:7.1f}  after {g[1]}  before {g[2]}")
    # exclusive time per kernel name: the time a kernel runs with NOTHING else in flight (it alone stretches the step)
    excl, tot, cnt = collections.Counter(), collections.Counter(), collections.Counter()
    for i, (s, e, n) in enumerate(ks):
        tot[n] += e - s
        cnt[n] += 1
    active = []
    events = sorted([(k[0], 0, i) for i, k in enumerate(ks)] + [(k[1], 1, i) for i, k in enumerate(ks)])
    live, last = set(), t0
    for t, kind, i in events:
        if len(live) == 1:
            excl[ks[next(iter(live))][2]] += t - last
        last = t
        if kind == 0:
            live.add(i)
        else:
            live.discard(i)
    P("# kernel name                                               n   total us   alone us")
    for n, v in tot.most_common(40):
        P(f"  {n:56s} {cnt[n]:4d} {v:9.1f} {excl[n]:9.1f}")
    # coarse phases: first / last kernel index of the backbone stages is unknown here; print a coarse time histogram instead
    P("# busy fraction per 100 us slice (sum of kernel time in slice / 100 us):")
    nb = int((t1 - t0) // 100) + 1
    sl = [0.0] * nb
    for s, e, n in ks:
        for k in range(max(0, int((s - t0) // 100)), min(nb - 1, int((e - t0) // 100)) + 1):      # integer slices: no float stepping
            lo, hi = t0 + 100.0 * k, t0 + 100.0 * (k + 1)
            sl[k] += max(0.0, min(e, hi) - max(s, lo))
    P("  " + " ".join(f"{v / 100:.1f}" for v in sl))
    if args.list:
        P("# ordered kernel list: start us, duration us, kernels in flight at start, name")
        for i, (s_, e_, n) in enumerate(ks):
            inflight = sum(1 for (a, b, _) in ks[max(0, i - 8):i] if b > s_)
            P(f"  {s_ - t0:8.1f} {e_ - s_:7.1f} {inflight:2d} {n}")
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
