"""bench.py — frames/sec of ASY-VRNet multi-task inference (512x512 RGB + 4x512x512 radar) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

ours      : EfficientVRNet(4, 9, phi) built from the vrcoc modules (hand-written sm_100a kernels behind the reference's
            nn.Module API), eval mode, bf16, batch 8 per GPU (BASELINE.json configs[2]: batch 64 sharded over 8 GPUs), the
            whole forward captured in one CUDA graph per pipeline slot.  One step = one forward over one batch of synthetic frames.
            `value`  = frames/s with the batch already resident in HBM (device-timed, max over ranks), two batches in flight
                       (vrcoc.InferenceSession: one graph + statistics arena per slot, two compute streams; every batch is
                       computed on its own);
            `serial` = the same with ONE graph replayed one batch at a time (= the per-batch latency);
            `e2e`    = frames/s through the public serving call with pinned-host inputs: H2D copy of the batch, forward, box
                       decode, D2H of the decoded boxes and of the per-pixel class map, all inside the timed region;
            `roofline` = the launch class with the largest share of the step, timed live with CUDA events per launch;
            `cpu_baseline` = the unmodified reference (baseline/_ref, vendored by __graft_entry__.build()) timed on this box's
                       host cores (rank 0, N=1); `reference_eager_gpu` = the same reference run eagerly on the GPU;
            `train_step` = a short BASELINE configs[3] measurement (the full one: --mode train).
reference : the unmodified reference's CPU forward of the same path (baseline/_ref; no product import in that arm), all host
            threads, one frame per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "asy-vrnet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "frames/sec (512x512 img+radar), ASY-VRNet multi-task inference"     # --img 1024 renames it in the line
UNIT = "frames/s"
PER_GPU_BATCH = 8
IMG = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "block-sweep"],
                    help="infer: BASELINE configs[2] (the driver's line); train: configs[3] (training step, NCCL gradient all-reduce); "
                         "block-sweep: configs[1] (tools/block_sweep.py; extra flags after --)")
    ap.add_argument("--img", type=int, default=512, help="frame side (1024: BASELINE configs[4], fea_pos buffers replaced)")
    ap.add_argument("--no-train-graph", action="store_true", help="train mode: eager forward/backward instead of the two CUDA graphs")
    ap.add_argument("--ref-loss", action="store_true", help="train mode: the reference's own per-image YOLOLoss instead of vrcoc.losses.YOLOLoss")
    ap.add_argument("--no-train", action="store_true", help="skip the short training-step measurement appended to the inference line")
    ap.add_argument("--no-check", action="store_true", help="skip the pre-timing oracle spot check of the benched outputs")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the unmodified reference eager on the GPU")
    ap.add_argument("--phi", default="l")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="frames per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="print the per-kernel time table to stderr")
    ap.add_argument("--ncu-pass", action="store_true",
                    help="profiling aid: run warmup+steps eager forwards (no CUDA graph, no timing, no JSON) and exit; used "
                         "under `ncu --metrics gpu__time_duration.sum` to produce profiles/*launches*.csv")
    args, rest = ap.parse_known_args()
    args.rest = [r for r in rest if r != "--"]
    if args.mode != "block-sweep" and args.rest:
        ap.error(f"unrecognized arguments: {args.rest}")
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe), in a thread for the duration of the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# per-kernel accounting: wrap the C-ABI entry points (python-side only; nothing changes in the library)
# ----------------------------------------------------------------------------------------------------------------
KERNELS_PER_CALL = {"vrcoc_channel_sums": 1, "vrcoc_conv_fwd": 1, "vrcoc_table_apply": 1, "vrcoc_cluster_core_fwd": 1,
                    "vrcoc_cluster_core_bwd": 2, "vrcoc_sa_gate_sums": 1, "vrcoc_radar_enh_table": 1, "vrcoc_chan_affine": 1,
                    "vrcoc_img_enh_finish": 1, "vrcoc_gelu_bwd": 1, "vrcoc_gn_bwd_sums": 1, "vrcoc_gn_bwd_apply": 1,
                    "vrcoc_conv1x1_wgrad": 3, "vrcoc_im2col": 1, "vrcoc_im2col_rows": 1, "vrcoc_patch_embed": 1, "vrcoc_upsample_bilinear": 1, "vrcoc_upsample_argmax": 1,
                    "vrcoc_dwconv": 1, "vrcoc_mlp_fused_fwd": 1, "vrcoc_token_mixer_fwd": 1, "vrcoc_token_mixer_core_fwd": 1}


def _esz(dt):
    return 4 if dt == 0 else 2


def _describe(name, args):
    """(label, algorithmic bytes, flops) of one call, from its arguments (SURVEY §8d per-unit figures)."""
    if name in ("vrcoc_conv_fwd", "vrcoc_table_apply"):
        d = args[0]._obj if hasattr(args[0], "_obj") else args[0]
        P_in, P_out, Cin = d.H_in * d.W_in, d.H_out * d.W_out, d.C0 + d.C1
        by = d.B * d.C0 * P_in * _esz(d.src0_dtype)
        if d.C1:
            by += (d.B if d.src1_bstride else 1) * d.C1 * P_in * _esz(d.src1_dtype)
        by += d.O * Cin * d.kh * d.kw * _esz(d.weight_dtype)
        by += d.B * d.O_split * P_out * _esz(d.out_dtype)
        if d.O_split != d.O:
            by += d.B * (d.O - d.O_split) * P_out * _esz(d.out2_dtype)
        if d.res:
            by += d.B * d.O * P_out * _esz(d.res_dtype)
        fl = 2.0 * d.B * d.O * Cin * d.kh * d.kw * P_out if name == "vrcoc_conv_fwd" else 0.0
        return f"conv{d.kh}x{d.kw}[{Cin}->{d.O}]@{d.H_out}x{d.W_out}", by, fl
    if name == "vrcoc_cluster_core_fwd":
        fdt, vdt, odt = args[1], args[3], args[5]
        B, E, D, H, W = args[10:15]
        pts = B * E * D * H * W
        M = args[17] * args[18]
        return f"cluster_core_fwd[E{E}xD{D}]@{H}x{W}", pts * (_esz(fdt) + _esz(vdt) + _esz(odt)), (2 * M + 5) * pts
    if name == "vrcoc_im2col":
        dt, B, C, H, W, kh, kw, stride, pad, dil = args[2:12]
        Ho, Wo = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1, (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        return f"im2col{kh}x{kw}[{C}]@{Ho}x{Wo}", B * C * _esz(dt) * (H * W + kh * kw * Ho * Wo), 0.0
    if name == "vrcoc_patch_embed":
        B, C0, C1, H, W, O = args[8:14]
        return (f"patch_embed4[{C0 + C1}->{O}]@{H // 4}x{W // 4}", 2 * (B * C0 * H * W + C1 * H * W + B * O * (H // 4) * (W // 4)),
                2.0 * B * O * (C0 + C1) * 16 * (H // 4) * (W // 4))
    if name == "vrcoc_im2col_rows":
        dt, B, C, H, W, kw, dil = args[2:9]
        return f"im2col_rows{kw}[{C}]@{H}x{W}", B * C * _esz(dt) * H * W * (1 + kw), 0.0
    if name == "vrcoc_channel_sums":
        dt, B, C, P = args[1:5]
        return f"channel_sums[{C}]@{P}pt", B * C * P * _esz(dt), 2.0 * B * C * P
    if name == "vrcoc_sa_gate_sums":
        dt, B, C, P = args[1:5]
        return f"sa_gate_sums[{C}]@{P}pt", B * C * P * _esz(dt), 8.0 * B * C * P
    if name == "vrcoc_mlp_fused_fwd":
        B, C, hid, P = args[12:16]
        side = int(round(P ** 0.5))
        return (f"mlp_fused[{C}->{hid}->{C}]@{side}x{P // side}", B * P * 3 * C * 2 + 4 * C * hid, 4.0 * B * P * C * hid)
    if name == "vrcoc_token_mixer_fwd":
        # fused token-mixer half: algorithmic bytes = x in + out out (2*C*P*s) + the folded weights once
        B, C, H, W, E, D = args[15:21]
        P, ED = H * W, E * D
        return (f"token_mixer_fused[{C}|{E}x{D}]@{H}x{W}", B * P * 2 * C * 2 + (2 * ED * 2 * C + C * ED) * 2,
                B * P * (2.0 * C * ED * 4 + (2 * 4 + 5) * ED))
    if name == "vrcoc_token_mixer_core_fwd":
        B, C, H, W, E, D = args[11:17]
        P, ED = H * W, E * D
        return (f"token_mixer_proj_core[{C}|{E}x{D}]@{H}x{W}", B * P * (C + ED) * 2 + 3 * ED * C * 2, B * P * (2.0 * C * ED * 3 + (2 * 4 + 5) * ED))
    if name == "vrcoc_dwconv":
        dt, B, C, H, W, k, stride, pad = args[4:12]
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        return f"dwconv{k}x{k}[{C}]@{Ho}x{Wo}", B * C * _esz(dt) * (H * W + Ho * Wo), 2.0 * B * C * k * k * Ho * Wo
    if name == "vrcoc_upsample_bilinear":
        dt, planes, H, W, Ho, Wo = args[2:8]
        return f"upsample[{H}x{W}->{Ho}x{Wo}]", planes * _esz(dt) * (H * W + Ho * Wo), 0.0
    if name == "vrcoc_upsample_argmax":
        dt, B, C, H, W, Ho, Wo = args[2:9]
        return f"upsample_argmax[{C}x{H}x{W}->{Ho}x{Wo}]", B * (C * _esz(dt) * H * W + Ho * Wo), 0.0
    return name.replace("vrcoc_", ""), 0, 0.0


class KernelAccounting:
    def __init__(self):
        from vrcoc import _lib
        self.lib = _lib.lib
        self.orig = {n: getattr(self.lib, n) for n in KERNELS_PER_CALL}
        self.mode = None
        self.count = 0
        self.records = []        # (label, bytes, flops, start_event, end_event)

    def _wrap(self, name):
        fn = self.orig[name]

        def call(*args):
            if self.mode == "count":
                self.count += KERNELS_PER_CALL[name]
                return fn(*args)
            label, by, fl = _describe(name, args)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            self.records.append((label, by, fl, s, e))
            return rc
        return call

    def start(self, mode):
        self.mode, self.count, self.records = mode, 0, []
        for n in self.orig:
            setattr(self.lib, n, self._wrap(n))

    def stop(self):
        for n, fn in self.orig.items():
            setattr(self.lib, n, fn)
        self.mode = None

    def table(self):
        torch.cuda.synchronize()
        agg = {}
        for label, by, fl, s, e in self.records:
            t = s.elapsed_time(e)
            a = agg.setdefault(label, [0, 0.0, 0, 0.0])
            a[0] += 1; a[1] += t; a[2] += by; a[3] += fl
        return agg


# ----------------------------------------------------------------------------------------------------------------
def build_model(phi, dtype, device, img=512):
    import vrcoc
    torch.manual_seed(0)
    m = vrcoc.EfficientVRNet(num_classes=4, num_seg_classes=9, phi=phi).eval()
    if img != 512:
        vrcoc.replace_pos_buffers(m, img)          # the reference's positional grid is a fixed 512x512 buffer
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        # the reference's training-time init (nets/yolo_training.py:482-500): conv weights N(0, 0.02), BN weight N(1, 0.02)
        for mod in m.modules():
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.Conv1d)):
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.02)
            elif isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(1 + torch.randn(mod.weight.shape, generator=g) * 0.02)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
    return m.to(device=device, dtype=dtype)


def synth_batch(batch, seed, dtype, img=IMG):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, img, img, generator=g).to(dtype)
    r = torch.rand(batch, 4, img, img, generator=g).to(dtype)
    return x, r


def _fit_pos(ref_model, img):
    """the reference model with its fea_pos buffers regenerated for `img` (same formula; oracle-side helper, no product import)"""
    if img == 512:
        return ref_model
    for m in ref_model.modules():
        if isinstance(getattr(m, "fea_pos", None), torch.Tensor) and "fea_pos" in m._buffers:
            rw = torch.arange(0, img, step=1) / (img - 1.0)
            pos = torch.stack(torch.meshgrid(rw, rw, indexing="ij"), dim=-1).float() - 0.5
            m._buffers["fea_pos"] = pos
            m._buffers["fea_pos_r"] = pos.clone()
    return ref_model


def load_reference():
    """The UNMODIFIED reference (vendored to git-ignored baseline/_ref/ by __graft_entry__.build()) through its own public
    API: nets/efficient_vrnet.py:EfficientVRNet + nets/yolo_training.py:weights_init (train.py:297-298).  No product module
    is imported on this path.  Returns None when the vendored tree is absent (then the oracle port stands in)."""
    from oracle import ref_shim
    if not ref_shim.available(ref_shim.VENDORED):
        return None
    ref_shim.install(ref_shim.VENDORED)
    import contextlib
    import io
    from nets.efficient_vrnet import EfficientVRNet as RefNet
    from nets.yolo_training import weights_init

    def make(phi, seed=0):
        torch.manual_seed(seed)
        m = RefNet(4, 9, phi)
        with contextlib.redirect_stdout(io.StringIO()):
            weights_init(m)
        return m.eval()
    return make


def _time_cpu_forward(fwd, x, r, n, warm):
    with torch.no_grad():
        for _ in range(warm):
            fwd(x, r)
        t0 = time.perf_counter()
        for _ in range(n):
            fwd(x, r)
        return time.perf_counter() - t0


def run_reference(args):
    """CPU arm: the reference's own EfficientVRNet (BASELINE.json configs[0]: batch 1, fp32, 512x512 RGB + 4x512x512 radar,
    random init) on all host threads, 1 frame / step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    x, r = synth_batch(1, 100, torch.float32, args.img)
    make = load_reference()
    if make is not None:
        model = _fit_pos(make(args.phi), args.img)
        fwd, kind = (lambda a, b: model(a, b)), "reference"
        what = "the reference's own nets/efficient_vrnet.py:EfficientVRNet, unmodified (baseline/_ref)"
    else:
        # only when the vendored reference is missing: the oracle port needs a state dict of the right shapes, and the product's
        # module tree (CPU tensors, no kernel runs) is the one place left to get it from
        from oracle import coc_oracle as O
        import vrcoc
        torch.manual_seed(0)
        sd = {k: v.float() for k, v in vrcoc.EfficientVRNet(4, 9, args.phi).state_dict().items()}
        fwd, kind = (lambda a, b: O.efficient_vrnet_forward(a, b, sd, args.phi)), "port"
        what = "oracle port of the reference (baseline/_ref absent)"
    dt = _time_cpu_forward(fwd, x, r, args.steps, max(1, min(args.warmup, 3)))
    fps = args.steps / dt
    sample = f"1 frame/step (one frame of the {args.batch}-frame per-GPU batch), fp32, phi={args.phi}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC.replace("512x512", f"{args.img}x{args.img}"), "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ASY-VRNet(phi={args.phi}) multi-task inference fwd, {args.img}x{args.img} RGB + 4x{args.img}x{args.img} radar (CPU: {what})"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(args, model):
    """Rank 0, N=1: the reference's own model (strict-loaded with the benched weights: the drop-in contract at work) timed on
    the host cores over a bounded sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    x, r = synth_batch(1, 100, torch.float32, args.img)
    make = load_reference()
    if make is not None:
        ref = _fit_pos(make(args.phi), args.img)
        ref.load_state_dict(sd, strict=True)
        fwd, kind = (lambda a, b: ref(a, b)), "reference"
    else:
        from oracle import coc_oracle as O
        fwd, kind = (lambda a, b: O.efficient_vrnet_forward(a, b, sd, args.phi)), "port"
    dt1 = _time_cpu_forward(fwd, x, r, 1, 1)
    n = max(4, min(64, int(12.0 / max(dt1, 1e-3))))          # ~12 s of host time
    dt = _time_cpu_forward(fwd, x, r, n, 0)
    return {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n} frames, batch 1, fp32, phi={args.phi}, after 2 warm-up frames"}


def reference_eager_gpu(args, model, dev):
    """Secondary baseline ('the honest kernel to beat'): the UNMODIFIED reference running eager on this B200 with
    the benched weights and batch — fp32 with TF32 off (the valid fp32 oracle setting), fp32 with TF32 on (PyTorch's
    default for cuDNN), and under bf16 autocast.  cuDNN/cuBLAS/ATen kernels; no product code."""
    make = load_reference()
    if make is None:
        return None
    ref = _fit_pos(make(args.phi), args.img)
    ref.load_state_dict({k: v.detach().float().cpu() for k, v in model.state_dict().items()}, strict=True)
    ref = ref.to(dev)
    x, r = (t.to(dev) for t in synth_batch(args.batch, 100, torch.float32, args.img))
    out = {}

    def run(name, tf32, autocast):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
        try:
            with torch.no_grad(), ctx:
                for _ in range(3):
                    ref(x, r)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 10
                s.record()
                for _ in range(n):
                    ref(x, r)
                e.record()
                torch.cuda.synchronize()
            out[name] = {"value": args.batch * n / (s.elapsed_time(e) / 1e3), "unit": UNIT, "batch": args.batch}
        except Exception as ex:       # e.g. out of memory for the [R,M,N,D] temporaries
            out[name] = {"error": str(ex)[:120]}
    run("fp32_tf32_off", False, False)
    run("fp32_tf32_on", True, False)
    run("bf16_autocast", True, True)
    del ref
    torch.cuda.empty_cache()
    return out


def oracle_spot_check(args, model, static_in, batch, graph, graph_outs):
    """Before any timing: the benched configuration (bf16 weights, batch B, CUDA graph replay, side-stream branches) must
    produce finite outputs that agree with the CPU oracle evaluated in fp64 on the SAME bf16-rounded weights and inputs
    (whole batch: data_normal couples the samples).  Gated: rel. L2 error of each detection map as the GRAPH REPLAY
    produced it and of the segmentation logits (eager pass of the same modules; the graph only keeps the class map)
    <= tol; the graph's per-pixel class map must agree with the oracle's arg-max on >= 99 % of the pixels."""
    from oracle import coc_oracle as O
    tol = 3e-2 if args.dtype == "bf16" else 2e-3
    sx, sr = static_in
    with torch.no_grad():
        sx.copy_(batch[0]); sr.copy_(batch[1])
        if graph is not None:
            graph.replay()
            det = [d.double().cpu() for d in graph_outs[0]]
            cls = graph_outs[1].cpu().long()
            _, seg_logits = model(sx, sr)
        else:
            det, seg_logits = model(sx, sr)
            det = [d.double().cpu() for d in det]
            cls = seg_logits.argmax(1).cpu().long()
        seg_logits = seg_logits.double().cpu()
        sd = {k: v.detach().double().cpu() for k, v in model.state_dict().items()}
        t0 = time.perf_counter()
        rdet, rseg = O.efficient_vrnet_forward(batch[0].double().cpu(), batch[1].double().cpu(), sd, args.phi)
        dt_oracle = time.perf_counter() - t0

    def rel(a, b):
        return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    errs = {f"det{i}": rel(a, b) for i, (a, b) in enumerate(zip(det, rdet))}
    errs["seg_logits"] = rel(seg_logits, rseg)
    finite = all(torch.isfinite(t).all().item() for t in det + [seg_logits])
    agree = (cls == rseg.argmax(1)).double().mean().item()
    ok = finite and max(errs.values()) <= tol and agree >= 0.99
    res = {"ok": ok, "finite": finite, "rel_err": errs, "seg_class_agreement": agree, "tol": tol, "through_cuda_graph": graph is not None,
           "oracle": f"oracle/coc_oracle.py fp64 on the {args.dtype}-rounded weights and inputs, batch {batch[0].shape[0]}, {dt_oracle:.1f} s"}
    if not ok:
        raise SystemExit("bench.py: the benched configuration fails its oracle spot check: " + json.dumps(res))
    return res


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: training step with the multitask loss, batch 16 / GPU, bf16, NCCL gradient all-reduce
# ----------------------------------------------------------------------------------------------------------------
class FusedEMA:
    """nets/yolo_training.py:449-475 (ModelEMA) with the same decay ramp, as two multi-tensor launches per step instead of a
    Python loop over ~900 state-dict entries."""

    def __init__(self, model, decay=0.9999, tau=2000):
        import copy
        import math
        self.ema = copy.deepcopy(model).eval()
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self.updates, self.decay = 0, (lambda x: decay * (1 - math.exp(-x / tau)))
        self.dst = [v for v in self.ema.state_dict().values() if v.dtype.is_floating_point and v.numel()]
        self.src = [v for v in model.state_dict().values() if v.dtype.is_floating_point and v.numel()]

    def update(self):
        self.updates += 1
        d = self.decay(self.updates)
        with torch.no_grad():
            torch._foreach_mul_(self.dst, d)
            torch._foreach_add_(self.dst, self.src, alpha=1 - d)


def train_setup(args, dev, world, batch):
    """model + optimizer + loss + synthetic batch exactly as train.py:297-473 / utils_fit.py:86-118 put them together:
    weights_init, SGD(momentum 0.937, nesterov) over the three parameter groups, YOLOLoss(num_classes=4) + 5 * (Focal + Dice),
    EMA; autocast in bf16 (the reference uses fp16 + GradScaler: bf16 needs no scaler)."""
    import contextlib
    import io
    import vrcoc
    from oracle import ref_shim
    torch.manual_seed(0)
    model = vrcoc.EfficientVRNet(num_classes=4, num_seg_classes=9, phi=args.phi)
    loss_kind = "surrogate (sum of squares; baseline/_ref absent)"
    yolo_loss = focal = dice = None
    if ref_shim.available(ref_shim.VENDORED):
        ref_shim.install(ref_shim.VENDORED)
        from nets.deeplabv3_training import Dice_loss, Focal_Loss
        from nets.yolo_training import YOLOLoss, weights_init
        with contextlib.redirect_stdout(io.StringIO()):
            weights_init(model)
        import nets.yolo_training as _yt
        # the reference calls torch.cuda.empty_cache() once per image inside its SimOTA loop (yolo_training.py:169): 16 times per
        # step it hands the caching allocator's whole pool back to the driver and every later allocation of the step pays
        # cudaMalloc.  Numerically a no-op; replaced by one here (the loss code itself is the reference's, unmodified).
        _yt.torch = type("_TorchNoEmptyCache", (), {"__getattr__": lambda self, k: getattr(torch, k),
                                                    "cuda": type("_Cuda", (), {"__getattr__": lambda self, k: (lambda: None) if k == "empty_cache" else getattr(torch.cuda, k)})()})()
        yolo_loss, focal, dice = YOLOLoss(4, True), Focal_Loss, Dice_loss      # fp16=True: its SimOTA cost leaves autocast (yolo_training.py:240-247)
        loss_kind = "reference YOLOLoss(4, fp16=True) + 5*(Focal_Loss + Dice_loss) (nets/yolo_training.py:60, nets/deeplabv3_training.py:22,41)"
        if not getattr(args, "ref_loss", False):
            # the product's detection loss: the reference's YOLOLoss arithmetic (decode, SimOTA dynamic-k assignment, IoU / objectness /
            # class terms) over the whole batch with static shapes and no host sync; tests/test_losses.py pins value and gradients
            # against the reference class.  --ref-loss runs the reference's own per-image loop instead.
            yolo_loss = vrcoc.losses.YOLOLoss(4, True)
            loss_kind = ("vrcoc.losses.YOLOLoss(4, fp16=True) [batched sync-free SimOTA, = reference nets/yolo_training.py:60 to 1e-5] "
                         "+ 5*(reference Focal_Loss + Dice_loss, nets/deeplabv3_training.py:22,41)")
    model = model.to(dev).train()
    for p in model.parameters():
        if p.numel() == 0:
            p.requires_grad_(False)          # the six zero-size tensors of radar_enhance_by_image1.image_attn: no gradient, no bucket
    pg0, pg1, pg2 = [], [], []
    for k, v in model.named_modules():       # train.py:460-467
        if hasattr(v, "bias") and isinstance(v.bias, torch.nn.Parameter) and v.bias.numel():
            pg2.append(v.bias)
        if isinstance(v, torch.nn.BatchNorm2d) or "bn" in k:
            pg0.append(v.weight)
        elif hasattr(v, "weight") and isinstance(v.weight, torch.nn.Parameter) and v.weight.numel():
            pg1.append(v.weight)
    opt = torch.optim.SGD(pg0, 1e-3, momentum=0.937, nesterov=True)
    opt.add_param_group({"params": pg1, "weight_decay": 5e-4})
    opt.add_param_group({"params": pg2})
    ema = FusedEMA(model)
    g = torch.Generator().manual_seed(7 + int(os.environ.get("RANK", "0")))
    B, S = batch, args.img
    images = torch.randn(B, 3, S, S, generator=g).to(dev)
    radars = torch.rand(B, 4, S, S, generator=g).to(dev)
    targets = []
    for _ in range(B):                         # 3 boxes / image, cx cy w h class in pixels (utils_fit.py:34 batch layout)
        cxcy = torch.rand(3, 2, generator=g) * (S - 112) + 56
        wh = torch.rand(3, 2, generator=g) * 80 + 16
        cls = torch.randint(0, 4, (3, 1), generator=g).float()
        targets.append(torch.cat([cxcy, wh, cls], 1).to(dev))
    pngs = torch.randint(0, 9, (B, S, S), generator=g).to(dev)
    seg_labels = torch.nn.functional.one_hot(pngs, 10).float()
    weights = torch.ones(9, device=dev)

    graphed = not getattr(args, "no_train_graph", False)
    if graphed:
        # forward and backward of the network as two CUDA graphs (torch.cuda.make_graphed_callables replaces model.forward): the eager
        # step is host-bound (~9 K launches); the loss (SimOTA: data-dependent shapes, host syncs) stays eager between the two graphs
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_), torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
            torch.cuda.make_graphed_callables(model, (images, radars), allow_unused_input=True)
        torch.cuda.current_stream().wait_stream(s_)
    net = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):                       # (documented recipe for DDP around graphed callables)
            net = DDP(model, device_ids=[dev.index], bucket_cap_mb=25, gradient_as_bucket_view=True)    # train.py:367 (DDP over NCCL)
        torch.cuda.current_stream().wait_stream(s_)
    fwd_call = net

    def step(timing=None):
        def mark(name, t0):
            if timing is not None:
                torch.cuda.synchronize()
                timing[name] = timing.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return time.perf_counter()
        t = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=not graphed):
            det, seg = fwd_call(images, radars)
            t = mark("forward_ms", t)
            if yolo_loss is not None:
                loss_seg = focal(seg, pngs, weights, num_classes=9) + dice(seg, seg_labels)
                loss_det = yolo_loss([d.float() for d in det], targets)
                loss = loss_det + 5 * loss_seg
            else:
                loss = sum(d.float().square().mean() for d in det) + seg.float().square().mean()
        t = mark("loss_ms", t)
        loss.backward()
        t = mark("backward_ms", t)
        opt.step()
        ema.update()
        mark("optimizer_ema_ms", t)
        return loss

    nparam = sum(p.numel() for p in model.parameters() if p.requires_grad)
    step.graphed = graphed
    return step, net, model, loss_kind, nparam


def measure_train(args, dev, world, dist, batch, steps, warmup):
    step, net, model, loss_kind, nparam = train_setup(args, dev, world, batch)
    loss = None
    for _ in range(warmup):
        loss = step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / steps
    out = {"value": batch * world / (ms_step / 1e3), "unit": "frames/s", "ms_per_step": ms_step, "batch_per_gpu": batch,
           "global_batch": batch * world, "steps": steps, "warmup": warmup, "dtype": "bf16 autocast, fp32 master weights",
           "cuda_graphs": "network forward and backward replayed as two CUDA graphs; loss, optimizer, EMA eager" if step.graphed else "eager",
           "loss": loss_kind, "final_loss": float(loss.detach().float().item()), "finite": bool(torch.isfinite(loss.detach()).item()),
           "optimizer": "SGD(momentum 0.937, nesterov), 3 parameter groups (train.py:460-473), fused multi-tensor EMA",
           "trainable_params": nparam, "grad_bytes_per_step": nparam * 4}
    timing = {}
    for _ in range(3):                       # phase split of a step (host-synchronised after every phase: sums to more than ms_per_step)
        step(timing)
    out["phases_ms_synced"] = {k: v / 3 for k, v in timing.items()}
    if dist is not None:
        # exposed all-reduce time: the same steps without gradient synchronisation (DDP no_sync), max over ranks
        with net.no_sync():
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            dist.barrier()
            s.record()
            for _ in range(steps):
                step()
            e.record()
            torch.cuda.synchronize()
        ms2 = torch.tensor([s.elapsed_time(e)], device=dev)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        out["ms_per_step_no_allreduce"] = ms2.item() / steps
        out["allreduce_exposed_ms"] = ms_step - ms2.item() / steps
        out["collective"] = f"DDP bucketed NCCL all-reduce of {nparam * 4 / 1e6:.0f} MB fp32 gradients per step, overlapped with backward"
    del step, net, model
    torch.cuda.empty_cache()
    return out


def run_train(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch if args.batch != PER_GPU_BATCH else 16
    with ClockSampler(local) as clk:
        res = measure_train(args, dev, world, dist, batch, args.steps if args.steps != 200 else 20, max(args.warmup, 3))
    if rank == 0:
        line = {"metric": f"training frames/sec ({args.img}x{args.img} img+radar), ASY-VRNet multitask step", "value": res["value"], "unit": "frames/s",
                "n_gpus": world, "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"ASY-VRNet(phi={args.phi}) training step (fwd + multitask loss + bwd + SGD + EMA), batch {batch}/GPU, "
                                       f"gradient all-reduce over NCCL" if world > 1 else f"ASY-VRNet(phi={args.phi}) training step, batch {batch}, 1 GPU"},
                "clocks": clk.summary(), "train": res}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = True
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model = build_model(args.phi, dtype, dev, args.img)
    B = args.batch

    # a ring of distinct device-resident batches so no step re-reads a warm input
    NBUF = 3
    host = [tuple(t.pin_memory() for t in synth_batch(B, 100 + rank * 10 + i, dtype, args.img)) for i in range(NBUF)]
    devb = [(x.to(dev), r.to(dev)) for x, r in host]
    sx, sr = torch.empty_like(devb[0][0]), torch.empty_like(devb[0][1])

    def forward(x, r):
        # detection maps + per-pixel class map; the neck's serving switch fuses the x4 logits upsample with the arg-max
        model.backbone.seg_class_map = True
        try:
            det, seg = model(x, r)
        finally:
            model.backbone.seg_class_map = False
        return det, (seg if seg.dtype == torch.uint8 else seg.argmax(dim=1).to(torch.uint8))

    if args.ncu_pass:
        # warm-up forwards (they also build the memoised derived weights: hundreds of one-off copy kernels) stay outside the capture:
        # run under `ncu --profile-from-start off`, the profiled range is the `steps` steady-state forwards
        with torch.no_grad():
            for i in range(max(1, args.warmup)):
                sx.copy_(devb[i % NBUF][0]); sr.copy_(devb[i % NBUF][1])
                forward(sx, sr)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            for i in range(args.steps):
                sx.copy_(devb[i % NBUF][0]); sr.copy_(devb[i % NBUF][1])
                forward(sx, sr)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return

    graph = None
    with torch.no_grad():
        sx.copy_(devb[0][0]); sr.copy_(devb[0][1])
        for _ in range(2):
            out = forward(sx, sr)
        torch.cuda.synchronize()
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = forward(sx, sr)
    det_out, seg_out = out
    host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in list(det_out) + [seg_out]]

    def step_resident(i):
        x, r = devb[i % NBUF]
        sx.copy_(x, non_blocking=True); sr.copy_(r, non_blocking=True)     # device->device staging into the graph's inputs
        if graph is not None:
            graph.replay()
        else:
            with torch.no_grad():
                o = forward(sx, sr)
            for d, s in zip(list(det_out) + [seg_out], list(o[0]) + [o[1]]):
                d.copy_(s)

    # End to end through the product's serving API (vrcoc.InferenceSession): two pipeline slots (input buffers + captured graph +
    # pinned output buffers each), the host->device copy of step i+1 and the device->host read of step i-1 on their own streams
    # under the forward of step i.  Every step copies its inputs from pinned host memory and reads its results (decoded boxes and
    # the per-pixel class map) back inside the timed region; the host holds the results of step i-1 before it queues step i+1.
    import vrcoc
    E2E_SLOTS = int(os.environ.get("VRCOC_BENCH_E2E_SLOTS", "3"))
    RES_SLOTS = int(os.environ.get("VRCOC_BENCH_RES_SLOTS", "2"))
    sess = vrcoc.InferenceSession(model, batch=B, img=args.img, slots=E2E_SLOTS, decode=True, cuda_graph=not args.no_graph)
    slots = sess.slots
    pending = [0]

    def step_e2e(i):
        x, r = host[i % NBUF]
        sess.submit(x, r)
        pending[0] += 1
        if pending[0] > E2E_SLOTS - 1:
            sess.collect()
            pending[0] -= 1

    def drain_e2e():
        while pending[0] > 0:
            sess.collect()                                                  # the last step's read-back ends inside the timed region
            pending[0] -= 1

    # Device-resident throughput with the same serving pipeline, minus the host copies: two slots, each its own captured graph,
    # compute stream and statistics arena, so the forwards of consecutive batches overlap (InferenceSession(concurrent=True)); the
    # batch is staged device->device into the slot's input buffers.  `serial` below is the one-batch-at-a-time replay of ONE graph.
    sess_res = vrcoc.InferenceSession(model, batch=B, img=args.img, slots=RES_SLOTS, decode=False, cuda_graph=not args.no_graph)

    def step_resident2(i):
        x, r = devb[i % NBUF]
        sess_res.submit(x, r, readback=False)

    def timed(step_fn, drain=None, streams=()):
        for i in range(args.warmup):
            step_fn(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        main = torch.cuda.current_stream()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for st in streams:
            st.wait_event(s)                                                # nothing of the timed steps starts before the start event
        for i in range(args.steps):
            step_fn(args.warmup + i)
        if drain is not None:
            drain()                                                         # the last step's read-back ends inside the timed region
        for st in streams:
            main.wait_stream(st)                                            # the end event follows every stream the steps ran on
        e.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    check = None
    if rank == 0 and not args.no_check:
        check = oracle_spot_check(args, model, (sx, sr), devb[0], graph, (det_out, seg_out))
    with ClockSampler(local) as clk:                  # started before the warm-up steps so that a 60 ms timed region has samples
        for i in range(20):
            step_resident(i)
        torch.cuda.synchronize()
        ms_serial = timed(step_resident)
        ms_res = timed(step_resident2, streams=sess_res.streams())
        ms_e2e = timed(step_e2e, drain_e2e, streams=sess.streams())
    frames = B * world * args.steps
    value = frames / (ms_res / 1e3)
    value_serial = frames / (ms_serial / 1e3)
    e2e = frames / (ms_e2e / 1e3)

    # launches per step (our kernels only) and the live per-kernel table — eager passes outside the timed region
    acct = KernelAccounting()
    with torch.no_grad():
        acct.start("count")
        forward(sx, sr)
        acct.stop()
        launches_per_step = acct.count
        # per-launch timing: one stream, and the stream is held while the host queues the whole eager forward, so that the
        # event pairs bracket kernels running back to back (an eager launch otherwise starts on an idle GPU and its pair
        # also measures ~5 us of launch latency; concurrent branches would measure each other)
        from vrcoc import ops as _ops
        pair_streams, _ops.PAIR_STREAMS = _ops.PAIR_STREAMS, False
        acct.start("time")
        for _ in range(3):
            torch.cuda._sleep(int(1.9e6 * 40))
            forward(sx, sr)
            torch.cuda.synchronize()
        acct.stop()
        _ops.PAIR_STREAMS = pair_streams
    table = acct.table()
    total_ms = sum(v[1] for v in table.values())
    pk = peaks()
    # the dominant kernel = largest share of the step among the launches with an algorithmic byte / flop count
    top_label, top = max(((k, v) for k, v in table.items() if v[2] or v[3]), key=lambda kv: kv[1][1])
    n, t_ms, by, fl = top
    hbm = by / (t_ms / 1e3) / 1e9
    tfl = fl / (t_ms / 1e3) / 1e12
    hbm_frac, tc_frac = hbm / pk["hbm_gbs"], tfl / pk["bf16_tflops"]
    if hbm_frac >= tc_frac:
        roof = {"bound": "hbm", "achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_frac}
    else:
        roof = {"bound": "tensor", "achieved": tfl, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tc_frac}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram__bytes_read+write per launch from `ncu --set full`
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{top_label}|B{B}|{args.dtype}")
    roof.update({"traffic": traffic, "kernel": top_label, "launches_timed": n, "avg_us": 1e3 * t_ms / n,
                 "share_of_native_time": t_ms / total_ms, "peak_source": pk["source"],
                 "hbm_frac": hbm_frac, "tensor_frac": tc_frac})
    # the HBM-bound kernel of the CoC block (the cluster core, ~1 flop/byte), reported beside the dominant one
    core = [(k, v) for k, v in table.items() if k.startswith("cluster_core_fwd")]
    roof_core = None
    if core:
        c_label, (c_n, c_ms, c_by, c_fl) = max(core, key=lambda kv: kv[1][2] / kv[1][0])     # the launch class moving the most bytes
        c_hbm = c_by / (c_ms / 1e3) / 1e9
        roof_core = {"bound": "hbm", "achieved": c_hbm, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": c_hbm / pk["hbm_gbs"],
                     "kernel": c_label, "launches_timed": c_n, "avg_us": 1e3 * c_ms / c_n,
                     "share_of_native_time": sum(v[1] for _, v in core) / total_ms, "peak_source": pk["source"]}
    if args.profile_kernels and rank == 0:
        for label, (cnt, t, b_, f_) in sorted(table.items(), key=lambda kv: -kv[1][1]):
            print(f"  {label:44s} n={cnt:4d} {t / cnt * 1e3:9.1f} us/launch  {b_ / (t / 1e3) / 1e9 if t else 0:8.1f} GB/s "
                  f"{f_ / (t / 1e3) / 1e12 if t else 0:7.1f} TF/s  share {t / total_ms:5.1%}", file=sys.stderr)

    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = sum(t.numel() * t.element_size() for t in slots[0]["host"])
    line = {
        "metric": METRIC.replace("512x512", f"{args.img}x{args.img}"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"ASY-VRNet(phi={args.phi}) multi-task inference fwd (3 det maps + seg class map), {args.img}x{args.img} RGB + "
                               f"4x{args.img}x{args.img} radar, random init, batch {B}/GPU (global {B * world}), batch-sharded, no collective",
                   "cuda_graph": graph is not None,
                   "batches_in_flight": "2: vrcoc.InferenceSession pipeline slots, one captured graph + compute stream + statistics arena "
                                        "each; every batch of 8 is still computed on its own (`serial` = one batch at a time)",
                   "l2": f"inputs rotate over {NBUF} distinct batches; per-step activation traffic >> 126 MB L2"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "api": f"vrcoc.InferenceSession(model, batch, slots={E2E_SLOTS}, decode=True).submit / .collect",
                "loop": f"{E2E_SLOTS} pipeline slots (buffers + captured graph each): H2D of the next step and D2H of the previous one (decoded boxes "
                        "+ class map) on copy streams, the forwards of two consecutive steps on the session's two compute streams"},
        "serial": {"value": value_serial, "unit": UNIT, "ms_per_step": ms_serial / args.steps,
                   "what": "one CUDA graph replayed one batch at a time on one stream (= the per-batch latency of the forward); "
                           "`value` overlaps the forwards of two consecutive batches"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clk.summary(),
        "roofline": roof,
        "roofline_coc_core": roof_core,
        "output_check": check,
    }
    # whole-step fractions on SURVEY 8d's per-frame figures for the 27 ClusterBlocks (bf16: 101 MB, 66.6 GFLOP per frame at phi='l',
    # 512x512): what fraction of the HBM / tensor peak the CoC path of the step amounts to
    if args.phi == "l" and args.img == 512 and args.dtype == "bf16":
        fps_gpu = value / world
        line["roofline_step"] = {"coc_alg_MB_per_frame": 100.7, "coc_GFLOP_per_frame": 66.6,
                                 "hbm_frac": 100.7e6 * fps_gpu / 1e9 / pk["hbm_gbs"], "tc_frac": 66.6e9 * fps_gpu / 1e12 / pk["bf16_tflops"]}
    if rank == 0 and world == 1:
        from tools import block_sweep
        if args.img == 512:
            with torch.no_grad():
                line["roofline_block"] = block_sweep.live_rows(B, dtype, iters=10)  # per live row: all launches of one ClusterBlock fwd
        if not args.no_ref_gpu:
            line["reference_eager_gpu"] = reference_eager_gpu(args, model, dev)
    if not args.no_train and args.img == 512:
        # BASELINE configs[3] beside the headline: a short training-step measurement at the same N (its gradient all-reduce is the
        # one collective of this repository); the full run is `bench.py --mode train`
        del graph, slots, sess, sess_res
        torch.cuda.empty_cache()
        try:
            tr = measure_train(args, dev, world, dist, 16, 8, 3)
        except Exception as ex:      # never lose the inference line to the secondary measurement
            tr = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        line["train_step"] = tr
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, model)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.mode == "block-sweep":
        from tools import block_sweep
        return block_sweep.main(args.rest)
    if args.mode == "train" and args.impl == "ours":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --mode train needs a CUDA device")
        return run_train(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the CoC + fusion path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
