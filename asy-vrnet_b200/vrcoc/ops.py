"""Tensor-level wrappers and autograd Functions over the C-ABI kernels.

PyTorch is plumbing here: it owns device memory and streams; every arithmetic pass over a feature map is one of the
hand-written kernels in csrc/.  Only O(B*C)-sized bookkeeping (GroupNorm backward coefficients, BatchNorm running
statistics) is done with torch ops on tiny tensors.
"""
import ctypes as C

import os

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, BF16, ENGINE_AUTO, F32, ConvDesc, check, lib

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise _lib.VrcocError(f"unsupported dtype {t.dtype}: the CoC path runs in float32 or bfloat16") from None


# ------------------------------------------------------------------------------------------------------------
# autocast
# ------------------------------------------------------------------------------------------------------------
def _amp_cast(a):
    """activations ([B,C,H,W]) and projection / convolution weights (ndim >= 2) -> bf16; vectors (biases, norm affines, layer
    scales, sim_alpha / sim_beta), fp64 statistics and non-floating arguments stay as they are (the kernels consume them in fp32)"""
    if isinstance(a, torch.Tensor) and a.is_cuda and a.is_floating_point() and a.ndim >= 2 and a.dtype in (torch.float32, torch.float16):
        return a.to(torch.bfloat16)
    return a


def amp_function(cls):
    """Class decorator for the autograd Functions of the native path: under torch.autocast (reference utils/utils_fit.py:86-109
    runs the model under autocast with fp16 by default, train.py:56) the kernels compute in bf16 — fp32 master weights and fp32 /
    fp16 activations are cast on the way in, the forward body runs with autocast off, and autograd casts the returned gradients
    back to the dtype of the original inputs.  The kernels have no fp16 storage type; fp16 autocast regions therefore get bf16
    activations back (same exponent range as fp32, so no GradScaler is needed for them)."""
    fwd = cls.forward

    def forward(ctx, *args):
        if torch.is_autocast_enabled("cuda"):
            args = tuple(_amp_cast(a) for a in args)
            with torch.autocast("cuda", enabled=False):
                return fwd(ctx, *args)
        return fwd(ctx, *args)

    cls.forward = staticmethod(forward)
    return cls


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VrcocError("vrcoc kernels need CUDA tensors on an sm_100 device (no CPU fallback exists)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_SIDE_STREAMS = {}
PAIR_STREAMS = os.environ.get("VRCOC_PAIR_STREAMS", "1") != "0"   # False / env 0: serialise the two modality branches (debug / A-B)


def run_pair(f_main, f_side):
    """Run two INDEPENDENT closures concurrently: f_main on the current stream, f_side on a per-device side stream, joined
    before returning.  The image and the radar halves of VRCoC (vr_coc.py:589-675) and the segmentation / detection halves of
    the neck (coc_fpn_dual.py:199-224) do not depend on each other between fusion points; at the deep stages one half fills
    less than half of the 148 SMs, so the halves are launched side by side (inside a CUDA-graph capture the fork/join becomes
    two parallel branches of the graph).  Only without autograd; the caller keeps the closures' inputs alive until the join,
    which is what makes cross-stream use of caching-allocator blocks safe."""
    main = torch.cuda.current_stream()
    if not PAIR_STREAMS or torch.is_grad_enabled():
        return f_main(), f_side()
    key = main.device.index
    side = _SIDE_STREAMS.get(key)
    if side is None:
        side = _SIDE_STREAMS[key] = torch.cuda.Stream(device=main.device)
    if side == main:            # nested call from inside a side branch: nothing left to overlap with
        return f_main(), f_side()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        b = f_side()
    a = f_main()
    main.wait_stream(side)
    return a, b


FUSED_MLP_MAX_C = int(os.environ.get("VRCOC_FUSED_MLP_MAX_C", "384"))   # the kernel's limit; "128" = stages 1-2 only (A/B switch)
FUSED_MLP = True             # set False to run the channel MLP as two GEMM launches (debug / A-B)


def mlp_fused_ok(x, hidden):
    B, C, H, W = x.shape
    # policy: the kernel covers C <= 384, but past 128 channels (stage 3: 64 point tiles, 1.6 MB of weights streamed per CTA)
    # it is bound by the per-SM weight stream and two launches spread over more SMs are faster (tools/microbench.py s3_mlpf)
    return FUSED_MLP and C <= FUSED_MLP_MAX_C and x.dtype == torch.bfloat16 and bool(lib.vrcoc_mlp_fused_supported(1, C, hidden, H * W))


def mlp_fused_fwd(x, sums, gamma, beta, eps, w1, b1, w2, b2, ls, out_sums):
    """out = x + ls * (W2 . gelu(W1 . GN(x) + b1) + b2) in one launch (csrc/mlp_fused.cuh; reference vr_coc.py:208-228, :270-275)"""
    B, C, H, W = x.shape
    hid = w1.shape[0]
    out = torch.empty_like(x)
    check(lib.vrcoc_mlp_fused_fwd(_ptr(x), _ptr(sums), _ptr(gamma), _ptr(beta), float(eps), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2),
                                  _ptr(ls), _ptr(out), _ptr(out_sums), B, C, hid, H * W, _stream()), "mlp_fused_fwd")
    return out


FUSED_TOKEN_MIXER = True     # set False to run the token-mixer half as three launches (debug / A-B)


def token_mixer_fused_ok(x, heads, head_dim, fold_w, fold_h, pw, ph):
    B, C, H, W = x.shape
    return FUSED_TOKEN_MIXER and x.dtype == torch.bfloat16 and \
        bool(lib.vrcoc_token_mixer_supported(BF16, C, H, W, heads, head_dim, fold_w, fold_h, pw, ph))


def token_mixer_fused_fwd(x, sums, eps, w_fold, k0, k1, alpha, beta, w2, b2, ls, out_sums, heads, head_dim, fold_w, fold_h, save_aux=False):
    """out = x + ls * (W2 . cluster_core(fc1(GN(x)), fc_v(GN(x))) + b2) in one launch (csrc/token_mixer_fused.cu;
    reference vr_coc.py:155-192 inside :264-267).  Returns (out, idx, sim_max)."""
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    idx = torch.empty(B, heads, H, W, device=x.device, dtype=torch.uint8) if save_aux else None
    smax = torch.empty(B, heads, H, W, device=x.device, dtype=torch.float32) if save_aux else None
    check(lib.vrcoc_token_mixer_fwd(_ptr(x), _ptr(sums), float(eps), _ptr(w_fold), _ptr(k0), _ptr(k1), _ptr(alpha), _ptr(beta), _ptr(w2),
                                    _ptr(b2), _ptr(ls), _ptr(out), _ptr(out_sums), _ptr(idx), _ptr(smax), B, C, H, W, heads, head_dim,
                                    fold_w, fold_h, _stream()), "token_mixer_fwd")
    return out, idx, smax


def token_mixer_core_ok(x, heads, head_dim, fold_w, fold_h, pw, ph):
    B, C, H, W = x.shape
    return FUSED_TOKEN_MIXER and x.dtype == torch.bfloat16 and B <= 256 and \
        bool(lib.vrcoc_token_mixer_core_supported(BF16, C, H, W, heads, head_dim, fold_w, fold_h, pw, ph))


def token_mixer_core_fwd(x, sums, eps, w_fold, k0, k1, alpha, beta, heads, head_dim, fold_w, fold_h, save_aux=False):
    """o = cluster_core(fc1(GN(x)), fc_v(GN(x))) in one launch for the wide stage (csrc/token_mixer_fused.cu, second kernel;
    reference vr_coc.py:155-190).  Returns (o [B, heads*head_dim, H, W], idx, sim_max)."""
    B, C, H, W = x.shape
    o = torch.empty(B, heads * head_dim, H, W, device=x.device, dtype=x.dtype)
    idx = torch.empty(B, heads, H, W, device=x.device, dtype=torch.uint8) if save_aux else None
    smax = torch.empty(B, heads, H, W, device=x.device, dtype=torch.float32) if save_aux else None
    check(lib.vrcoc_token_mixer_core_fwd(_ptr(x), _ptr(sums), float(eps), _ptr(w_fold), _ptr(k0), _ptr(k1), _ptr(alpha), _ptr(beta), _ptr(o),
                                         _ptr(idx), _ptr(smax), B, C, H, W, heads, head_dim, fold_w, fold_h, _stream()), "token_mixer_core_fwd")
    return o, idx, smax


class Fork:
    """f() on side stream number `lane` (>= 1) of the current device, ordered after everything already queued on the
    current stream; join() orders the current stream after it and returns f's result.  Same liveness rule as run_pair: the
    object keeps the closure (and so its inputs) alive until the join.  Without CUDA / with autograd it degenerates to a
    plain call."""

    def __init__(self, f, lane):
        main = torch.cuda.current_stream() if torch.cuda.is_available() else None
        self.side = None
        if main is not None and PAIR_STREAMS and not torch.is_grad_enabled():
            key = (main.device.index, lane)
            side = _SIDE_STREAMS.get(key)
            if side is None:
                side = _SIDE_STREAMS[key] = torch.cuda.Stream(device=main.device)
            if side != main:
                self.side = side
        self.f = f
        if self.side is None:
            self.value = f()
        else:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                self.value = f()

    def join(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self.f = None
        return self.value


def _f32(t):
    """Small parameter vectors are always consumed as fp32.  The converted copy is memoised on the tensor object and
    keyed on (storage pointer, version counter): optimizer steps / load_state_dict / .to() invalidate it, so inference
    pays the conversion kernel once instead of once per call."""
    if t is None:
        return None
    if t.dtype == torch.float32 and t.is_contiguous():
        return t.detach()
    sig = (t.data_ptr(), t._version)
    memo = t.__dict__.get("_vrcoc_f32") if hasattr(t, "__dict__") else None
    if memo is not None and memo[0] == sig:
        return memo[1]
    if memo is not None and memo[1].shape == t.shape and memo[1].device == t.device:
        # refreshed IN PLACE: a captured CUDA graph that reads the old copy keeps reading valid, current memory
        with torch.no_grad():
            memo[1].copy_(t.detach())
        v = memo[1]
    else:
        v = t.detach().float().contiguous()
    try:
        t._vrcoc_f32 = (sig, v)
    except Exception:
        pass
    return v


def cached(mod, key, sources, build):
    """Derived parameter tensors (concatenated projection weights, folded BatchNorm affines, ...) memoised on the module
    while their sources are unchanged.  Bypassed whenever autograd could need the sources."""
    if torch.is_grad_enabled() and any(s is not None and s.requires_grad for s in sources):
        return build()
    store = mod.__dict__.setdefault("_vrcoc_cache", {})
    sig = tuple(None if s is None else (s.data_ptr(), s._version, s.dtype) for s in sources)
    ent = store.get(key)
    if ent is None or ent[0] != sig:
        with torch.no_grad():
            new = build()
            if ent is not None and _refresh_in_place(ent[1], new):
                ent = (sig, ent[1])          # same storage, new values: pointers baked into captured CUDA graphs stay valid
            else:
                ent = (sig, new)
        store[key] = ent
    return ent[1]


def _refresh_in_place(old, new):
    """copy `new` into the storage of `old` when both are tensors (or equal-length tuples of tensors) of the same shape / dtype /
    device.  Graph-capture contract: buffers whose pointers were handed to kernels are never freed or replaced while
    their module lives; derived parameter tensors are refreshed in place after load_state_dict / optimizer steps."""
    if isinstance(old, torch.Tensor) and isinstance(new, torch.Tensor):
        if old.shape == new.shape and old.dtype == new.dtype and old.device == new.device:
            old.copy_(new)
            return True
        return False
    if isinstance(old, (tuple, list)) and isinstance(new, (tuple, list)) and len(old) == len(new):
        if all(isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.shape == b.shape and a.dtype == b.dtype and a.device == b.device
               for a, b in zip(old, new)):
            for a, b in zip(old, new):
                a.copy_(b)
            return True
    return False


# ------------------------------------------------------------------------------------------------------------
# statistics
# ------------------------------------------------------------------------------------------------------------
def channel_sums(x, want_chan=True, want_sample=False):
    """-> (chan_sums float [B,C,2] or None, sample_sums double [B,STAT_SLOTS,2] or None)"""
    _need_cuda(x)
    x = x.contiguous()
    B, Cc, H, W = x.shape
    cs = torch.empty(B, Cc, 2, device=x.device, dtype=torch.float32) if want_chan else None
    ss = new_sample_sums(B, x.device) if want_sample else None
    check(lib.vrcoc_channel_sums(_ptr(x), _dt(x), B, Cc, H * W, _ptr(cs), _ptr(ss), _stream()), "channel_sums")
    return cs, ss


STAT_SLOTS = 32      # == VRCOC_STAT_SLOTS (include/vrcoc.h)


class sums_arena:
    """Context manager around one gradient-free model forward: all statistics buffers of the forward come out of ONE
    device buffer that is zeroed once on entry (one fill kernel instead of ~40 per forward; they sit on the critical path of
    the launch chain).  Slices carry the arena generation; a statistics tensor that rode along on an activation from an
    earlier forward is recognised as stale by sample_sums_of and recomputed.  Nested uses share the outermost arena."""
    _state = {}      # (device index, B) -> [buffer, next_free, generation, depth]; one arena per batch size, never freed:
                     # a CUDA graph captured at batch size B keeps zeroing / accumulating into ITS arena (graph-capture contract)
    _open = {}       # device index -> key of the arena opened by the outermost context
    SLOTS = 128
    lane = 0         # forwards that may run CONCURRENTLY on one device (the pipeline slots of InferenceSession: one captured
                     # graph each, replayed on their own streams) must not share an arena: each is captured under its own lane

    _lanes = 0

    @staticmethod
    def new_lane():
        """a lane id no other forward uses (lane 0 = everything that runs on the caller's stream)"""
        sums_arena._lanes += 1
        return sums_arena._lanes

    def __init__(self, B, device):
        self.key = None
        if device.type == "cuda" and not torch.is_grad_enabled():
            self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
            self.key, self.B, self.device = (self.dev_index, B, sums_arena.lane), B, device

    def __enter__(self):
        if self.key is None:
            return self
        cur = sums_arena._open.get(self.dev_index)
        if cur is not None and cur != self.key:
            self.key = None              # a different batch size inside an open arena: leave it alone
            return self
        st = sums_arena._state.get(self.key)
        if st is None:
            st = sums_arena._state[self.key] = [torch.empty(sums_arena.SLOTS, self.B, STAT_SLOTS, 2, device=self.device, dtype=torch.float64), 0, 0, 0]
        if st[3] == 0:
            st[0].zero_()
            st[1] = 0
            st[2] += 1
            sums_arena._open[self.dev_index] = self.key
        st[3] += 1
        return self

    def __exit__(self, *exc):
        if self.key is not None:
            st = sums_arena._state[self.key]
            st[3] -= 1
            if st[3] == 0:
                sums_arena._open.pop(self.dev_index, None)
        return False

    @staticmethod
    def take(B, device, n):
        if device.type != "cuda" or torch.is_grad_enabled():
            return None
        di = device.index if device.index is not None else torch.cuda.current_device()
        st = sums_arena._state.get(sums_arena._open.get(di))
        k = 1 if n is None else n
        if st is None or st[3] == 0 or st[0].shape[1] != B or st[1] + k > sums_arena.SLOTS:
            return None
        out = st[0][st[1]:st[1] + k] if n is not None else st[0][st[1]]
        st[1] += k
        out._vrcoc_arena = (st[0], st[2])
        return out

    @staticmethod
    def stale(ss):
        tag = getattr(ss, "_vrcoc_arena", None)
        if tag is None:
            return False
        for st in sums_arena._state.values():
            if st[0] is tag[0]:
                return st[2] != tag[1]
        return True


def tag_like(view, parent):
    """carry the arena tag of `parent` over to a view of it"""
    tag = getattr(parent, "_vrcoc_arena", None)
    if tag is not None:
        view._vrcoc_arena = tag
    return view


def new_sample_sums(B, device, n=None):
    """zeroed slot-wise GroupNorm statistics buffer(s): [B, STAT_SLOTS, 2] (or [n, B, STAT_SLOTS, 2]) float64"""
    device = torch.device(device)
    got = sums_arena.take(B, device, n)
    if got is not None:
        return got
    shape = (B, STAT_SLOTS, 2) if n is None else (n, B, STAT_SLOTS, 2)
    return torch.zeros(shape, device=device, dtype=torch.float64)


def attach_sums(x, ss):
    """let the per-sample statistics of x ride along on the tensor, tagged with its version counter and storage pointer: an
    in-place update of x between two blocks (y.mul_(s), nn.ReLU(inplace=True), ...) invalidates them"""
    x._vrcoc_sums = ss
    x._vrcoc_sums_tag = (x._version, x.data_ptr())
    return x


def sample_sums_of(x):
    """GroupNorm(1,C) statistics of x: reuse the producer's side output when it rode along on the tensor."""
    ss = getattr(x, "_vrcoc_sums", None)
    tag = getattr(x, "_vrcoc_sums_tag", None)
    if ss is not None and ss.shape[0] == x.shape[0] and not sums_arena.stale(ss) and (tag is None or tag == (x._version, x.data_ptr())):
        return ss
    return channel_sums(x, want_chan=False, want_sample=True)[1]


# ------------------------------------------------------------------------------------------------------------
# conv engine
# ------------------------------------------------------------------------------------------------------------
def conv_desc(src0, weight, out, *, src1=None, src0_bstride=None, src1_bstride=None, chan_src=None,
              gn=None, table=None, has_gate=False, kh=1, kw=1, stride=1, pad=0,
              e_scale=None, e_shift=None, act=ACT_NONE, post_scale=None, res=None, f_scale=None, f_shift=None,
              out2=None, out_sample_sums=None, out_minmax=None, engine=ENGINE_AUTO, keep=None, dil=1, k_order=0, gn_fold=None):
    """Fill a ConvDesc from tensors.  `keep` collects temporaries that must outlive the launch call."""
    d = ConvDesc()
    B = out.shape[0]
    C0 = src0.shape[-3]
    H_in, W_in = src0.shape[-2], src0.shape[-1]
    C1 = 0 if src1 is None else src1.shape[-3]
    d.B, d.H_in, d.W_in, d.H_out, d.W_out = B, H_in, W_in, out.shape[2], out.shape[3]
    d.C0, d.C1 = C0, C1
    O_split = out.shape[1]
    d.O = O_split + (0 if out2 is None else out2.shape[1])
    d.kh, d.kw, d.stride, d.pad = kh, kw, stride, pad
    d.src0, d.src0_dtype = _ptr(src0), _dt(src0)
    d.src0_bstride = C0 * H_in * W_in if src0_bstride is None else src0_bstride
    if src1 is not None:
        d.src1, d.src1_dtype = _ptr(src1), _dt(src1)
        d.src1_bstride = C1 * H_in * W_in if src1_bstride is None else src1_bstride
    d.chan_src = _ptr(chan_src)
    if gn_fold is not None:
        # GroupNorm folded into the weights: `weight` is [O][2*C0] = [hi | lo], e_shift = k0, statistics enter in the epilogue
        sums, k1, eps = gn_fold
        d.gn_sums, d.gn_eps, d.gn_fold_k1 = _ptr(sums), eps, _ptr(k1)
    elif gn is not None:
        sums, gamma, beta, eps = gn
        d.gn_sums, d.gn_gamma, d.gn_beta, d.gn_eps = _ptr(sums), _ptr(gamma), _ptr(beta), eps
    d.table, d.has_gate = _ptr(table), int(has_gate)
    d.weight, d.weight_dtype = _ptr(weight), _dt(weight)
    d.e_scale, d.e_shift, d.act, d.post_scale = _ptr(e_scale), _ptr(e_shift), act, _ptr(post_scale)
    if res is not None:
        d.res, d.res_dtype = _ptr(res), _dt(res)
    d.f_scale, d.f_shift = _ptr(f_scale), _ptr(f_shift)
    d.out, d.out_dtype = _ptr(out), _dt(out)
    if out2 is not None:
        d.out2, d.out2_dtype = _ptr(out2), _dt(out2)
    d.O_split = O_split
    d.out_sample_sums, d.out_minmax = _ptr(out_sample_sums), _ptr(out_minmax)
    d.engine = engine
    d.dil, d.k_order = dil, k_order
    return d


FOLD_GN = True              # set False to run GN -> fc1|fc_v with the in-place prologue (bf16-rounded GN(x) operand; debug / A-B)


def gn_fold_ok(x, O, O_split):
    """Can the GN -> fc1|fc_v projection of this activation run in the folded form (include/vrcoc.h, gn_fold_k1)?"""
    B, C, H, W = x.shape
    return FOLD_GN and x.dtype == torch.bfloat16 and bool(lib.vrcoc_gn_fold_supported(B, C, O, O_split, H * W))


def fold_gn_weights(w_feat, b_feat, w_value, b_value, gamma, beta):
    """GroupNorm(1,C) folded into the concatenated fc1 | fc_v projection (reference vr_coc.py:156-157 after :265):

        W.GN(x) + b = rstd * ((W diag(gamma)) . x) - rstd * mean * k1 + k0,   k1 = (W diag(gamma)) . 1,  k0 = b + W . beta

    so the tensor core contracts the RAW bf16 activations and nothing is rounded before the similarity: W diag(gamma) is
    split into two bf16 terms hi + lo (error 2^-17).  Rounding GN(x) itself to bf16 — the only alternative for a bf16
    operand — flips ~0.05 % of the arg-max assignments and moves the Cluster output by ~2e-2.
    Returns (w_fold [2ED][2C] bf16, k0 [2ED] fp32, k1 [2ED] fp32).  The fc_v rows carry a lo half only when the feat | value
    boundary is not on a 128-row tile boundary (neck: ED = 96), where every tile contracts both halves; otherwise `value`
    (stored as bf16 anyway) is computed from the hi half alone and its lo half is zero."""
    w = torch.cat([w_feat, w_value], 0).float().reshape(w_feat.shape[0] + w_value.shape[0], -1)      # [2ED][C], bf16-exact values
    wg = w * gamma.float()[None, :]
    hi = wg.to(torch.bfloat16)
    lo = (wg - hi.float()).to(torch.bfloat16)
    ED = w_feat.shape[0]
    if ED % 128 == 0:
        lo[ED:] = 0                                   # value rows contract the hi half only (their tiles stop after C)
    k1 = (hi.double() + lo.double()).sum(1).float()
    k0 = (torch.cat([b_feat, b_value], 0).double() + (w.double() * beta.double()[None, :]).sum(1)).float()
    return torch.cat([hi, lo], 1).contiguous(), k0.contiguous(), k1.contiguous()


def tap_major(weight):
    """[O, C, kh, kw] -> [O, kh*kw*C] (k = tap*C + c), memoised on the weight tensor like `_f32`.  With this K order the
    im2col rows of one 64-wide k slab are consecutive channels of a single tap, which makes the tcgen05 gather cheap."""
    sig = (weight.data_ptr(), weight._version)
    memo = weight.__dict__.get("_vrcoc_tapmajor") if hasattr(weight, "__dict__") else None
    if memo is not None and memo[0] == sig:
        return memo[1]
    v = weight.detach().permute(0, 2, 3, 1).reshape(weight.shape[0], -1).contiguous()
    if memo is not None and memo[1].shape == v.shape and memo[1].dtype == v.dtype and memo[1].device == v.device:
        with torch.no_grad():
            memo[1].copy_(v)                  # in place: see _refresh_in_place
        v = memo[1]
    try:
        weight._vrcoc_tapmajor = (sig, v)
    except Exception:
        pass
    return v


def center_tap(weight):
    """[O, C, k, k] -> [O, C]: the centre tap as a contiguous 1x1 weight, memoised on the weight tensor like `tap_major`.  A dilated
    k x k convolution whose dilation is >= the map size reads only zero padding through every other tap (ASPP's rate-18 branch on
    the 16x16 map, reference neck/coc_fpn_dual.py:55-67)."""
    sig = (weight.data_ptr(), weight._version)
    memo = weight.__dict__.get("_vrcoc_centertap") if hasattr(weight, "__dict__") else None
    if memo is not None and memo[0] == sig:
        return memo[1]
    v = weight.detach()[:, :, weight.shape[2] // 2, weight.shape[3] // 2].contiguous()
    if memo is not None and memo[1].shape == v.shape and memo[1].dtype == v.dtype and memo[1].device == v.device:
        with torch.no_grad():
            memo[1].copy_(v)                  # in place: see _refresh_in_place
        v = memo[1]
    try:
        weight._vrcoc_centertap = (sig, v)
    except Exception:
        pass
    return v


def conv_fwd(desc):
    check(lib.vrcoc_conv_fwd(C.byref(desc), _stream()), "conv_fwd")


def conv1x1_wgrad(desc, dy, want_db=True):
    """dW [O,Cin] fp32, db [O] fp32 for the forward call described by `desc`."""
    O, Cin = desc.O, desc.C0 + desc.C1
    n = lib.vrcoc_conv1x1_wgrad_workspace(C.byref(desc))
    if n < 0:
        check(-1, "conv1x1_wgrad_workspace")
    ws = torch.empty(n, device=dy.device, dtype=torch.float32)
    dW = torch.empty(O, Cin, device=dy.device, dtype=torch.float32)
    db = torch.empty(O, device=dy.device, dtype=torch.float32) if want_db else None
    check(lib.vrcoc_conv1x1_wgrad(C.byref(desc), _ptr(dy), _dt(dy), _ptr(dW), _ptr(db), _ptr(ws), n, _stream()), "conv1x1_wgrad")
    return dW, db


def out_hw(h, w, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1


# ------------------------------------------------------------------------------------------------------------
# cluster core
# ------------------------------------------------------------------------------------------------------------
def _bstride(t):
    """batch stride (elements) of a [B, C, H, W] tensor that is dense in its last three dims"""
    B, Cc, H, W = t.shape
    if t.stride(3) != 1 or t.stride(2) != W or t.stride(1) != H * W:
        raise _lib.VrcocError("tensor must be dense in (C,H,W)")
    return t.stride(0) if B > 1 else Cc * H * W


def _chw_dense(t):
    B, Cc, H, W = t.shape
    return t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W


def cluster_core_fwd(feat, value, alpha, beta, heads, fold_w, fold_h, proposal_w, proposal_h, out_dtype=None,
                     save_aux=False):
    _need_cuda(feat, value)
    if not _chw_dense(feat):
        feat = feat.contiguous()
    if not _chw_dense(value):
        value = value.contiguous()
    B, ED, H, W = feat.shape
    if ED % heads:
        raise _lib.VrcocError(f"channels {ED} not divisible by heads {heads}")
    D = ED // heads
    out = torch.empty(B, ED, H, W, device=feat.device, dtype=out_dtype or value.dtype)
    idx = torch.empty(B, heads, H, W, device=feat.device, dtype=torch.uint8) if save_aux else None
    smax = torch.empty(B, heads, H, W, device=feat.device, dtype=torch.float32) if save_aux else None
    check(lib.vrcoc_cluster_core_fwd(_ptr(feat), _dt(feat), _ptr(value), _dt(value), _ptr(out), _dt(out), _ptr(idx), _ptr(smax),
                                     _ptr(alpha), _ptr(beta), B, heads, D, H, W, fold_w, fold_h, proposal_w, proposal_h,
                                     _bstride(feat), _bstride(value), 0, _stream()), "cluster_core_fwd")
    return out, idx, smax


def cluster_core_bwd(feat, value, dout, idx, smax, alpha, beta, heads, fold_w, fold_h, proposal_w, proposal_h,
                     dfeat=None, dvalue=None):
    B, ED, H, W = feat.shape
    D = ED // heads
    if not _chw_dense(dout):
        dout = dout.contiguous()
    if dfeat is None:
        dfeat = torch.empty(B, ED, H, W, device=feat.device, dtype=feat.dtype)
    if dvalue is None:
        dvalue = torch.empty(B, ED, H, W, device=feat.device, dtype=value.dtype)
    f1, f2 = (fold_w, fold_h) if (fold_w > 1 and fold_h > 1) else (1, 1)
    R = B * heads * f1 * f2
    partials = torch.empty(2 * R, device=feat.device, dtype=torch.float32)
    dab = torch.empty(2, device=feat.device, dtype=torch.float32)
    check(lib.vrcoc_cluster_core_bwd(_ptr(feat), _dt(feat), _ptr(value), _dt(value), _ptr(dout), _dt(dout), _ptr(idx), _ptr(smax),
                                     _ptr(alpha), _ptr(beta), _ptr(dfeat), _dt(dfeat), _ptr(dvalue), _dt(dvalue), _ptr(dab),
                                     _ptr(partials), B, heads, D, H, W, fold_w, fold_h, proposal_w, proposal_h,
                                     _bstride(feat), _bstride(value), _bstride(dout), _bstride(dfeat), _bstride(dvalue),
                                     _stream()), "cluster_core_bwd")
    return dfeat, dvalue, dab


class ClusterCoreFn(torch.autograd.Function):          # (no autocast cast: `feat` must stay fp32)
    """out = cluster_core(feat, value; alpha, beta)   (reference vr_coc.py:158-190)"""

    @staticmethod
    def forward(ctx, feat, value, alpha, beta, heads, fold_w, fold_h, proposal_w, proposal_h):
        a32, b32 = _f32(alpha), _f32(beta)
        need_grad = any(ctx.needs_input_grad[:4])
        out, idx, smax = cluster_core_fwd(feat, value, a32, b32, heads, fold_w, fold_h, proposal_w, proposal_h,
                                          save_aux=need_grad)
        if need_grad:
            ctx.save_for_backward(feat, value, idx, smax, a32, b32)
            ctx.cfg = (heads, fold_w, fold_h, proposal_w, proposal_h)
            ctx.ab_dtype = alpha.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        feat, value, idx, smax, a32, b32 = ctx.saved_tensors
        # feat|value may be two views of one projection output: write the gradients into one buffer of the same
        # layout so the projection backward consumes a single dense dy without a concat pass
        B, ED, H, W = feat.shape
        joint = (feat.dtype == value.dtype and feat.stride(0) == 2 * ED * H * W and value.stride(0) == feat.stride(0)
                 and value.data_ptr() == feat.data_ptr() + ED * H * W * feat.element_size())
        if joint:
            buf = torch.empty(B, 2 * ED, H, W, device=feat.device, dtype=feat.dtype)
            dfeat, dvalue = buf[:, :ED], buf[:, ED:]
        else:
            dfeat = dvalue = None
        dfeat, dvalue, dab = cluster_core_bwd(feat, value, dout, idx, smax, a32, b32, *ctx.cfg, dfeat=dfeat, dvalue=dvalue)
        return dfeat, dvalue, dab[0:1], dab[1:2], None, None, None, None, None


# ------------------------------------------------------------------------------------------------------------
# projections (1x1 convs) with GroupNorm prologue / layer-scale + residual epilogue
# ------------------------------------------------------------------------------------------------------------
def _w2d(w):
    return w.detach().reshape(w.shape[0], -1)


@amp_function
class GNProjFn(torch.autograd.Function):
    """y = act(W * GroupNorm1(x) + b), optionally split along the output channels into (y[:split] as fp32, rest).
    `sums=None` skips the GroupNorm prologue (stand-alone Cluster.forward).

    Replaces norm -> 1x1 conv (-> GELU) of reference vr_coc.py:156-157 / :218-219 with one kernel; the normalised
    activation never reaches HBM."""

    @staticmethod
    def forward(ctx, x, sums, gamma, beta, eps, weight, bias, act, split_fp32):
        _need_cuda(x)
        x = x.contiguous()
        B, Cc, H, W = x.shape
        w2 = _w2d(weight).contiguous()
        O = w2.shape[0]
        g32, b32, bias32 = _f32(gamma), _f32(beta), _f32(bias)
        if split_fp32 and x.dtype != torch.float32:
            out = torch.empty(B, split_fp32, H, W, device=x.device, dtype=torch.float32)
            out2 = torch.empty(B, O - split_fp32, H, W, device=x.device, dtype=x.dtype)
        else:
            out = torch.empty(B, O, H, W, device=x.device, dtype=x.dtype)
            out2 = None
        gn = None if sums is None else (sums, g32, b32, eps)
        if gn is not None and out2 is not None and act == ACT_NONE and w2.dtype == torch.bfloat16 and gn_fold_ok(x, O, split_fp32):
            # fc1 | fc_v after norm1: GroupNorm folded into split-bf16 weights, `feat` exact to 2^-17 (see fold_gn_weights)
            bz = bias32 if bias32 is not None else torch.zeros(O, device=x.device)
            w_fold, k0, k1 = fold_gn_weights(w2[:split_fp32], bz[:split_fp32], w2[split_fp32:], bz[split_fp32:], g32, b32)
            d = conv_desc(x, w_fold, out, gn_fold=(sums, k1, eps), e_shift=k0, out2=out2)
        else:
            d = conv_desc(x, w2, out, gn=gn, e_shift=bias32, act=act, out2=out2)
        conv_fwd(d)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, sums, g32, b32, w2, bias32)
            ctx.meta = (eps, act, None if gamma is None else gamma.dtype, weight.shape, weight.dtype,
                        None if bias is None else bias.dtype, out2 is not None)
        if out2 is not None:
            return out, out2
        return out

    @staticmethod
    def backward(ctx, *grads):
        x, sums, g32, b32, w2, bias32 = ctx.saved_tensors
        eps, act, gdt, wshape, wdt, bdt, was_split = ctx.meta
        B, Cc, H, W = x.shape
        P = H * W
        O = w2.shape[0]
        if was_split:
            dy = torch.cat([grads[0].to(x.dtype), grads[1].to(x.dtype)], dim=1)
        else:
            dy = grads[0]
        dy = dy.contiguous()
        gn = None if sums is None else (sums, g32, b32, eps)
        if act == ACT_GELU:
            # recompute the pre-activation instead of having saved the (r*C x P) hidden map
            u = torch.empty(B, O, H, W, device=x.device, dtype=x.dtype)
            conv_fwd(conv_desc(x, w2, u, gn=gn, e_shift=bias32))
            check(lib.vrcoc_gelu_bwd(_ptr(dy), _ptr(u), _ptr(u), _dt(u), u.numel(), _stream()), "gelu_bwd")
            dy = u
        elif act != ACT_NONE:
            raise _lib.VrcocError("GNProjFn backward supports act none / gelu")
        # weight / bias gradients:  dW = dy . GN(x)^T
        fdesc = conv_desc(x, w2, dy, gn=gn)
        dW, db = conv1x1_wgrad(fdesc, dy, want_db=bias32 is not None)
        # input gradient through the projection: dz = W^T dy
        wt = w2.t().contiguous()
        dz = torch.empty(B, Cc, H, W, device=x.device, dtype=x.dtype)
        conv_fwd(conv_desc(dy, wt, dz))
        if gn is None:
            return (dz, None, None, None, None, dW.reshape(wshape), db, None, None)
        # GroupNorm(1,C) backward: per-(b,c) sums -> coefficients (one small kernel pair) -> dx = dz*a + x*bb + cc
        s = torch.empty(B, Cc, 2, device=x.device, dtype=torch.float32)
        check(lib.vrcoc_gn_bwd_sums(_ptr(dz), _ptr(x), _dt(x), B, Cc, P, _ptr(s), _stream()), "gn_bwd_sums")
        small = torch.empty(2 * B * Cc + 2 * B + 2 * Cc, device=x.device, dtype=torch.float32)
        a, ws = small[:B * Cc], small[B * Cc:2 * B * Cc]
        o = 2 * B * Cc
        bb, cc, dgamma, dbeta = small[o:o + B], small[o + B:o + 2 * B], small[o + 2 * B:o + 2 * B + Cc], small[o + 2 * B + Cc:]
        check(lib.vrcoc_gn_bwd_coef(_ptr(s), _ptr(sums), _ptr(g32), float(eps), B, Cc, P, _ptr(a), _ptr(bb), _ptr(cc), _ptr(dgamma),
                                    _ptr(dbeta), _ptr(ws), _stream()), "gn_bwd_coef")
        dx = torch.empty_like(x)
        check(lib.vrcoc_gn_bwd_apply(_ptr(dz), _ptr(x), None, _ptr(dx), _dt(x), _ptr(a), _ptr(bb), _ptr(cc), B, Cc, P, _stream()),
              "gn_bwd_apply")
        # gradients are returned in fp32: autograd casts them to the dtype of the corresponding input
        return (dx, None, dgamma, dbeta, None, dW.reshape(wshape), db, None, None)


@amp_function
class ProjResidualFn(torch.autograd.Function):
    """out = res + ls * (W h + b)  (+ per-sample {sum, sum^2} of out for the next GroupNorm as a side output).

    Replaces 1x1 conv -> layer-scale -> residual add of reference vr_coc.py:191,222 and :266-271."""

    @staticmethod
    def forward(ctx, h, weight, bias, ls, res):
        _need_cuda(h, res)
        h = h.contiguous()
        res = res.contiguous()
        B, K, H, W = h.shape
        w2 = _w2d(weight).contiguous()
        O = w2.shape[0]
        bias32, ls32 = _f32(bias), _f32(ls)
        out = torch.empty(B, O, H, W, device=h.device, dtype=res.dtype)
        sums = new_sample_sums(B, h.device)
        conv_fwd(conv_desc(h, w2, out, e_shift=bias32, post_scale=ls32, res=res, out_sample_sums=sums))
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(h, w2, bias32, ls32)
            ctx.meta = (weight.shape, weight.dtype, None if bias is None else bias.dtype, None if ls is None else ls.dtype)
        ctx.mark_non_differentiable(sums)
        return out, sums

    @staticmethod
    def backward(ctx, dout, _dsums):
        h, w2, bias32, ls32 = ctx.saved_tensors
        wshape, wdt, bdt, lsdt = ctx.meta
        dout = dout.contiguous()
        B, K, H, W = h.shape
        # G[o,k] = sum dout[o,p] h[k,p];  s[o] = sum dout[o,p]
        G, s = conv1x1_wgrad(conv_desc(h, w2, dout), dout, want_db=True)
        O = w2.shape[0]
        small = torch.empty(O * K + 2 * O, device=h.device, dtype=torch.float32)
        dW, db, dls = small[:O * K].view(O, K), small[O * K:O * K + O], small[O * K + O:]
        wt = torch.empty(K, O, device=h.device, dtype=w2.dtype)
        check(lib.vrcoc_proj_res_bwd_coef(_ptr(G), _ptr(s), _ptr(w2), _dt(w2), _ptr(bias32), _ptr(ls32), O, K, _ptr(dW), _ptr(db), _ptr(dls),
                                          _ptr(wt), _stream()), "proj_res_bwd_coef")
        dh = torch.empty_like(h)
        conv_fwd(conv_desc(dout, wt, dh))
        return (dh, dW.reshape(wshape), None if bias32 is None else db, None if ls32 is None else dls, dout)


@amp_function
class ProjFn(torch.autograd.Function):
    """y = act(W x + b) for a 1x1 projection (stand-alone Cluster / Mlp forward without the block fusion)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        _need_cuda(x)
        x = x.contiguous()
        B, K, H, W = x.shape
        w2 = _w2d(weight).contiguous()
        bias32 = _f32(bias)
        out = torch.empty(B, w2.shape[0], H, W, device=x.device, dtype=x.dtype)
        conv_fwd(conv_desc(x, w2, out, e_shift=bias32, act=act))
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, w2, bias32)
            ctx.meta = (act, weight.shape, weight.dtype, None if bias is None else bias.dtype)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w2, bias32 = ctx.saved_tensors
        act, wshape, wdt, bdt = ctx.meta
        dy = dy.contiguous()
        if act == ACT_GELU:
            u = torch.empty_like(dy)
            conv_fwd(conv_desc(x, w2, u, e_shift=bias32))
            check(lib.vrcoc_gelu_bwd(_ptr(dy), _ptr(u), _ptr(u), _dt(u), u.numel(), _stream()), "gelu_bwd")
            dy = u
        elif act != ACT_NONE:
            raise _lib.VrcocError("ProjFn backward supports act none / gelu")
        dW, db = conv1x1_wgrad(conv_desc(x, w2, dy), dy, want_db=bias32 is not None)
        dx = torch.empty_like(x)
        conv_fwd(conv_desc(dy, w2.t().contiguous(), dx))
        return dx, dW.reshape(wshape), db, None
