// C-ABI entry points of the convolution-as-GEMM engine: argument validation, engine selection, and the
// weight-gradient kernels of the 1x1 projections.
#include "conv_common.cuh"

namespace vrcoc {

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int esize(int dt) { return dt == VRCOC_F32 ? 4 : 2; }

static int fill_args(const vrcoc_conv_desc* d, ConvArgs& a) {
  VRCOC_REQUIRE(d != nullptr, "conv: null descriptor");
  VRCOC_REQUIRE(d->B > 0 && d->H_in > 0 && d->W_in > 0 && d->H_out > 0 && d->W_out > 0 && d->O > 0 && d->C0 > 0 && d->C1 >= 0,
                "conv: non-positive dimension");
  VRCOC_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad >= 0, "conv: bad kernel geometry");
  const int dil = d->dil > 0 ? d->dil : 1;
  const bool rowtap = d->k_order == 2;
  if (rowtap) {
    // vertical taps only (the horizontal ones are materialised in src0): padding applies to the rows, the map size is kept
    VRCOC_REQUIRE(d->kw == 1 && (d->kh & 1) == 1 && d->stride == 1 && d->pad == dil * (d->kh - 1) / 2 && d->H_out == d->H_in &&
                      d->W_out == d->W_in && d->C1 == 0 && !d->chan_src && !d->gn_sums && !d->table && !d->has_gate && !d->gn_fold_k1 &&
                      d->src0_dtype == VRCOC_BF16 && d->weight_dtype == VRCOC_BF16 && (d->C0 % 64) == 0 && ((dil * d->W_in) % 8) == 0,
                  "conv: row-tap mode (k_order 2) needs kw = 1, stride 1, pad = dil * (kh - 1) / 2, one bf16 source with C0 %% 64 == 0, "
                  "dil * W %% 8 == 0 and no prologue");
  }
  VRCOC_REQUIRE(rowtap || (d->H_out == (d->H_in + 2 * d->pad - dil * (d->kh - 1) - 1) / d->stride + 1 &&
                    d->W_out == (d->W_in + 2 * d->pad - dil * (d->kw - 1) - 1) / d->stride + 1),
                "conv: output size %dx%d inconsistent with input %dx%d k=%dx%d s=%d p=%d", d->H_out, d->W_out, d->H_in, d->W_in,
                d->kh, d->kw, d->stride, d->pad);
  VRCOC_REQUIRE(d->src0 && d->weight && d->out, "conv: null src0/weight/out");
  VRCOC_REQUIRE(d->C1 == 0 || d->src1, "conv: C1 > 0 but src1 is null");
  VRCOC_REQUIRE(d->O_split > 0 && d->O_split <= d->O && (d->O_split == d->O || d->out2), "conv: bad O_split / out2");
  VRCOC_REQUIRE(!(d->gn_sums && d->table), "conv: gn_sums and table are mutually exclusive");
  VRCOC_REQUIRE(!d->gn_sums || d->gn_fold_k1 || (d->gn_gamma && d->gn_beta && d->C1 == 0 && !d->chan_src),
                "conv: GroupNorm prologue needs gamma/beta and a single source");
  VRCOC_REQUIRE(!d->gn_fold_k1 || (d->gn_sums && !d->gn_gamma && !d->gn_beta && !d->e_scale && d->e_shift && d->C1 == 0 && !d->chan_src &&
                                   !d->table && d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0 && (d->C0 % 64) == 0 &&
                                   d->weight_dtype == VRCOC_BF16 && d->src0_dtype == VRCOC_BF16),
                "conv: folded GroupNorm needs gn_sums, e_shift (k0), a bf16 1x1 projection of a single bf16 source with C0 %% 64 == 0");
  auto okdt = [](int t) { return t == VRCOC_F32 || t == VRCOC_BF16; };
  VRCOC_REQUIRE(okdt(d->src0_dtype) && okdt(d->weight_dtype) && okdt(d->out_dtype) && (d->C1 == 0 || okdt(d->src1_dtype)) &&
                    (!d->res || okdt(d->res_dtype)) && (d->O_split == d->O || okdt(d->out2_dtype)),
                "conv: unknown dtype");
  a.B = d->B; a.H_in = d->H_in; a.W_in = d->W_in; a.H_out = d->H_out; a.W_out = d->W_out;
  a.C0 = d->C0; a.C1 = d->C1; a.Cin = d->C0 + d->C1; a.O = d->O;
  a.kh = d->kh; a.kw = d->kw; a.stride = d->stride; a.pad = d->pad; a.dil = dil; a.k_order = d->k_order ? 1 : 0;
  a.K = a.Cin * d->kh * d->kw;
  a.P_in = d->H_in * d->W_in; a.P_out = d->H_out * d->W_out;
  a.src0 = d->src0; a.src0_dtype = d->src0_dtype; a.src0_bstride = d->src0_bstride;
  a.src1 = d->src1; a.src1_dtype = d->src1_dtype; a.src1_bstride = d->src1_bstride;
  a.chan_src = d->chan_src;
  a.gn_sums = d->gn_sums; a.gn_gamma = d->gn_gamma; a.gn_beta = d->gn_beta; a.gn_eps = d->gn_eps;
  a.gn_fold_k1 = d->gn_fold_k1;
  a.table = d->table; a.has_gate = d->has_gate;
  a.weight = d->weight; a.weight_dtype = d->weight_dtype;
  a.e_scale = d->e_scale; a.e_shift = d->e_shift; a.act = d->act; a.post_scale = d->post_scale;
  a.res = d->res; a.res_dtype = d->res_dtype; a.f_scale = d->f_scale; a.f_shift = d->f_shift;
  a.out = d->out; a.out_dtype = d->out_dtype; a.out2 = d->out2; a.out2_dtype = d->out2_dtype; a.O_split = d->O_split;
  a.out_sample_sums = d->out_sample_sums; a.out_minmax = d->out_minmax;
  bool one = d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0;
  bool src_al = aligned16(d->src0) && (d->src0_bstride * esize(d->src0_dtype)) % 16 == 0 &&
                (d->C1 == 0 || (aligned16(d->src1) && (d->src1_bstride * esize(d->src1_dtype)) % 16 == 0));
  a.fast1x1 = one && a.P_in % 8 == 0 && src_al;
  a.rt_taps = a.rt_C = a.rt_dil = a.rt_W = 0;
  if (rowtap) {
    // from here on a 1x1 projection over kh * C0 virtual channels; only the TMA kernels understand the virtual source
    VRCOC_REQUIRE(src_al && a.P_in % 8 == 0, "conv: row-tap mode needs a 16-byte aligned source and H * W %% 8 == 0");
    a.rt_taps = d->kh; a.rt_C = d->C0; a.rt_dil = dil; a.rt_W = d->W_in;
    a.C0 = a.Cin = a.K = d->C0 * d->kh;
    a.kh = a.kw = 1; a.pad = 0; a.dil = 1; a.k_order = 0;
    a.fast1x1 = 1;
  }
  a.vec_out = a.P_out % 8 == 0 && aligned16(d->out) && (d->O_split == d->O || aligned16(d->out2)) && (!d->res || aligned16(d->res));
  return VRCOC_OK;
}

// ---- weight gradient of a 1x1 projection ------------------------------------------------------------------------
//   dW[o][k] = sum_{b,p} dy[b,o,p] * z[b,k,p],  z = prologue(src)     db[o] = sum_{b,p} dy[b,o,p]
// Split over (b, point-slab) units; each CTA reduces its units for one 64x64 tile of dW into the workspace, a second
// kernel adds the splits in a fixed order (deterministic, no float atomics).
constexpr int WG_T = 64, WG_P = 32, WG_RS = 68, WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS)
wgrad_kernel(ConvArgs a, const void* __restrict__ dy, int dy_dtype, float* __restrict__ ws, float* __restrict__ ws_db,
             int units_per_split, int slabs_per_sample) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* Ys = reinterpret_cast<float*>(smem);          // [WG_P][WG_RS]  dy slab, point-major
  float* Zs = Ys + WG_P * WG_RS;                       // [WG_P][WG_RS]  z slab
  float4* tab = reinterpret_cast<float4*>(Zs + WG_P * WG_RS);
  const int tid = threadIdx.x;
  const int o0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T, split = blockIdx.z;
  const int row = tid & 63, pseg = (tid >> 6) * 8;
  const int to = tid & 15, tk = tid >> 4;
  float acc[4][4];
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int P = a.P_out;
  const int total_units = a.B * slabs_per_sample;
  const int u0 = split * units_per_split;
  const int u1 = min(total_units, u0 + units_per_split);
  int cur_b = -1;
  for (int u = u0; u < u1; ++u) {
    const int b = u / slabs_per_sample;
    const int pbase = (u - b * slabs_per_sample) * WG_P;
    if (b != cur_b) {          // uniform across the CTA
      __syncthreads();
      build_prologue_table(a, b, tab);
      cur_b = b;
    }
    __syncthreads();
    // dy slab
    {
      const int o = o0 + row;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int p = pbase + pseg + i;
        float v = (o < a.O && p < P) ? ld_any(dy, ((int64_t)b * a.O + o) * P + p, dy_dtype) : 0.f;
        Ys[(pseg + i) * WG_RS + row] = v;
      }
    }
    // z slab (prologue applied)
    {
      const int k = k0 + row;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      const void* src = a.src0; int dt = a.src0_dtype; int64_t base = 0;
      if (k < a.Cin) {
        t = tab[k];
        const int s = a.chan_src ? a.chan_src[k] : k;
        if (s < a.C0) { base = (int64_t)b * a.src0_bstride + (int64_t)s * P; }
        else { src = a.src1; dt = a.src1_dtype; base = (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * P; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int p = pbase + pseg + i;
        float y = 0.f;
        if (k < a.Cin && p < P) {
          float x = ld_any(src, base + p, dt);
          y = fmaf(x, t.x, t.y);
          if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, x, t.w));
        }
        Zs[(pseg + i) * WG_RS + row] = y;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < WG_P; ++p) {
      float4 yv = *reinterpret_cast<const float4*>(Ys + p * WG_RS + to * 4);
      float4 zv = *reinterpret_cast<const float4*>(Zs + p * WG_RS + tk * 4);
      const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
      const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yy[i], zz[j], acc[i][j]);
        dbacc[i] += yy[i];
      }
    }
  }
  float* w = ws + (int64_t)split * a.O * a.Cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int o = o0 + to * 4 + i;
    if (o >= a.O) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tk * 4 + j;
      if (k < a.Cin) w[(int64_t)o * a.Cin + k] = acc[i][j];
    }
    if (ws_db && blockIdx.y == 0 && tk == 0) ws_db[(int64_t)split * a.O + o] = dbacc[i];
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += ws[(int64_t)k * n + i];
  out[i] = s;
}

// ---- stand-alone prologue application ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_apply_kernel(ConvArgs a) {
  const int b = blockIdx.y, k = blockIdx.x;
  const int P = a.P_in;
  float4 t;
  if (a.gn_sums) {
    float mu, rstd;
    gn_mean_rstd(a.gn_sums, b, (double)a.C0 * (double)P, a.gn_eps, mu, rstd);
    float sc = rstd * a.gn_gamma[k];
    t = make_float4(sc, fmaf(-mu, sc, a.gn_beta[k]), 0.f, 88.f);
  } else if (a.table) {
    t = reinterpret_cast<const float4*>(a.table)[(int64_t)b * a.Cin + k];
  } else {
    t = make_float4(1.f, 0.f, 0.f, 88.f);
  }
  const int s = a.chan_src ? a.chan_src[k] : k;
  const void* src = a.src0; int dt = a.src0_dtype; int64_t base;
  if (s < a.C0) base = (int64_t)b * a.src0_bstride + (int64_t)s * P;
  else { src = a.src1; dt = a.src1_dtype; base = (int64_t)b * a.src1_bstride + (int64_t)(s - a.C0) * P; }
  const int64_t obase = ((int64_t)b * a.Cin + k) * P;
  if (dt == VRCOC_BF16 && a.out_dtype == VRCOC_BF16 && (P & 7) == 0 && ((base | obase) & 7) == 0 &&
      ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(a.out)) & 15) == 0) {
    // bf16 planes (the neck's cat + shuffle copies, the attention gates): 128-bit loads / stores, 8 points per thread
    const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(src) + base;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + obase;
    const bool ident = t.x == 1.f && t.y == 0.f && !a.has_gate;
    for (int i = threadIdx.x * 8; i < P; i += blockDim.x * 8) {
      if (ident) {
        *reinterpret_cast<uint4*>(op + i) = __ldg(reinterpret_cast<const uint4*>(sp + i));
      } else {
        float v[8];
        ld8<__nv_bfloat16>(sp + i, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = fmaf(v[e], t.x, t.y);
          if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, v[e], t.w));
          v[e] = y;
        }
        st8<__nv_bfloat16>(op + i, v);
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    float x = ld_any(src, base + i, dt);
    float y = fmaf(x, t.x, t.y);
    if (a.has_gate) y *= sigmoidf_exact(fmaf(t.z, x, t.w));
    st_any(a.out, obase + i, a.out_dtype, y);
  }
}

int64_t wgrad_tc_workspace(const ConvArgs& a, int dy_dtype);                                   // wgrad_tc.cu
int launch_wgrad_tc(const ConvArgs& a, const void* dy, int dy_dtype, float* dW, float* db, float* ws, cudaStream_t st);

static int wgrad_splits(const ConvArgs& a, int& slabs_per_sample, int& units_per_split) {
  slabs_per_sample = (int)cdiv(a.P_out, WG_P);
  int units = a.B * slabs_per_sample;
  int tiles = (int)(cdiv(a.O, WG_T) * cdiv(a.Cin, WG_T));
  int want = (int)cdiv(148 * 4, tiles);
  int splits = units < want ? units : want;
  if (splits < 1) splits = 1;
  units_per_split = (int)cdiv(units, splits);
  return (int)cdiv(units, units_per_split);
}

}  // namespace vrcoc

using namespace vrcoc;

extern "C" int vrcoc_conv_fwd(const vrcoc_conv_desc* d, void* stream) {
  ConvArgs a;
  int rc = fill_args(d, a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.gn_fold_k1) return launch_conv_tc(a, st);        // only the channel-major tcgen05 kernel implements the fold
  if (a.rt_taps) {                                       // row-tap mode: shifted TMA boxes, tcgen05 kernels only
    VRCOC_REQUIRE(conv_tc_supported(a), "conv: row-tap mode: the problem is not supported by the tcgen05 engine");
    return launch_conv_tc(a, st);
  }
  if (d->engine == 2) {
    VRCOC_REQUIRE(conv_tc_supported(a), "conv: tcgen05 engine forced but the problem is not supported by it");
    return launch_conv_tc(a, st);
  }
  if (d->engine == 0 && conv_small_supported(a)) return launch_conv_small(a, st);
  if (d->engine == 0 && conv_tc_supported(a)) return launch_conv_tc(a, st);
  return launch_conv_simt(a, st);
}

extern "C" int vrcoc_table_apply(const vrcoc_conv_desc* d, void* stream) {
  VRCOC_REQUIRE(d != nullptr, "table_apply: null descriptor");
  vrcoc_conv_desc dd = *d;
  dd.weight = d->src0;              // not used; keeps the shared validation happy
  dd.weight_dtype = d->src0_dtype;
  ConvArgs a;
  int rc = fill_args(&dd, a);
  if (rc) return rc;
  VRCOC_REQUIRE(a.O == a.Cin && a.kh == 1 && a.kw == 1 && a.stride == 1 && a.pad == 0 && a.O_split == a.O && !a.rt_taps,
                "table_apply: needs O == C0+C1 and a 1x1 geometry");
  dim3 grid((unsigned)a.Cin, (unsigned)a.B);
  table_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("table_apply");
}

extern "C" int64_t vrcoc_conv1x1_wgrad_workspace(const vrcoc_conv_desc* d) {
  ConvArgs a;
  if (fill_args(d, a) || a.rt_taps) return -1;
  int sps, ups;
  int splits = wgrad_splits(a, sps, ups);
  const int64_t simt = (int64_t)splits * ((int64_t)a.O * a.Cin + a.O);
  // the descriptor's `out` stands for dy (same shape and dtype): enough for whichever path vrcoc_conv1x1_wgrad takes
  const int64_t tc = wgrad_tc_workspace(a, a.out_dtype);
  return tc > simt ? tc : simt;
}

extern "C" int vrcoc_conv1x1_wgrad(const vrcoc_conv_desc* d, const void* dy, int dy_dtype, float* dW, float* db,
                                   float* workspace, int64_t workspace_floats, void* stream) {
  ConvArgs a;
  int rc = fill_args(d, a);
  if (rc) return rc;
  VRCOC_REQUIRE(dy && dW && workspace, "wgrad: null pointer");
  VRCOC_REQUIRE(a.kh == 1 && a.kw == 1 && a.stride == 1 && a.pad == 0 && !a.rt_taps, "wgrad: only 1x1 projections are supported");
  int sps, ups;
  int splits = wgrad_splits(a, sps, ups);
  int64_t need = (int64_t)splits * ((int64_t)a.O * a.Cin + a.O);
  VRCOC_REQUIRE(workspace_floats >= need, "wgrad: workspace too small (%lld < %lld floats)", (long long)workspace_floats,
                (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  // bf16 operands: tensor-core path on the raw activations, prologue applied to the per-sample partials (wgrad_tc.cu)
  const int64_t tc_need = wgrad_tc_workspace(a, dy_dtype);
  if (tc_need >= 0 && workspace_floats >= tc_need) {
    rc = launch_wgrad_tc(a, dy, dy_dtype, dW, db, workspace, st);
    if (rc != 1) return rc;
  }
  float* ws = workspace;
  float* ws_db = workspace + (int64_t)splits * a.O * a.Cin;
  size_t smem = (size_t)2 * WG_P * WG_RS * sizeof(float) + (size_t)a.Cin * sizeof(float4);
  VRCOC_REQUIRE(smem <= 200 * 1024, "wgrad: too many input channels");
  if (smem > 48 * 1024) cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)cdiv(a.O, WG_T), (unsigned)cdiv(a.Cin, WG_T), (unsigned)splits);
  wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(a, dy, dy_dtype, ws, db ? ws_db : nullptr, ups, sps);
  rc = check_launch("wgrad");
  if (rc) return rc;
  int64_t n = (int64_t)a.O * a.Cin;
  wgrad_reduce_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(ws, splits, n, dW);
  if (db) wgrad_reduce_kernel<<<(unsigned)cdiv(a.O, 256), 256, 0, st>>>(ws_db, splits, a.O, db);
  return check_launch("wgrad.reduce");
}
