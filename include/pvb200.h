/*
 * pvb200.h -- C ABI of libpvb200.so: the B200 (sm_100a) kernels behind the Conv3d PV-yield step.
 *
 * This is the drop-in boundary for ONE hot path of openclimatefix/predict_pv_yield: the Conv3d
 * model's train / inference step.  The reference is pure Python + torch and has no FFI of its own
 * (SURVEY.md section 2a); every entry point below therefore replaces a torch operator call site in the
 * reference, cited as file:line relative to the reference root.  A maintainer binds these with
 * ctypes; predict_pv_yield_b200/lib.py is exactly that binding.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes, caller-owned workspaces, a cudaStream_t passed
 *     as void*.  No allocation, no retained pointers, no implicit synchronisation inside.
 *   - every function returns 0 (PVB200_OK) or a non-zero status; pvb200_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - tensors are dense, row-major, in torch's native layouts: activations NCDHW
 *     ([B][C][T][H][W]), Conv3d weights [Cout][Cin][3][3][3], Linear weights [out][in].
 *   - fp32 functions end in _f32; bf16 tensor-core functions end in _bf16 (bf16 storage as
 *     uint16_t, fp32 accumulation).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PVB200_H
#define PVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVB200_ABI_VERSION 1

#define PVB200_OK 0
#define PVB200_ERR_INVALID 1   /* bad argument / unsupported shape */
#define PVB200_ERR_CUDA 2      /* CUDA runtime error (message has cudaGetErrorString) */
#define PVB200_ERR_WORKSPACE 3 /* workspace too small */

typedef void* pvb200_stream_t; /* cudaStream_t */

/* ---- library state ------------------------------------------------------------------------ */
int pvb200_abi_version(void);
const char* pvb200_last_error(void);
/* number of kernels this library has launched in this process (bench.py's "gpu_launches") */
unsigned long long pvb200_launch_count(void);
void pvb200_reset_launch_count(void);
/* SM count of the current device (grid sizing), or <0 on error */
/* Under data parallelism NCCL's kernels occupy some SMs while the convolution backward runs; persistent kernels with
 * one CTA per SM then wait for the displaced CTAs.  pvb200_reserve_sms(n) makes every persistent kernel launched
 * afterwards size its grid for (SM count - n) SMs (n = 0: default).  Returns the previous value. */
int pvb200_reserve_sms(int n);
/* The other answer to the same problem: persistent kernels that support it (the fp32-mode weight gradient) stop splitting
 * their work statically over the CTAs and claim it in chunks from an atomic counter, so that CTAs displaced by NCCL's kernels
 * cost a chunk, not a grid tail.  Off by default (on one GPU the static split is perfectly balanced; chunks cost a few per
 * cent of tail); predict_pv_yield_b200/dp.py switches it on when the world size is > 1.  Returns the previous setting. */
int pvb200_set_dynamic_tiles(int on);
int pvb200_sm_count(void);
/* diagnostic: launch an FP32 FMA saturation kernel; *flops_out = FLOPs it performs.  bench.py times it
 * with CUDA events to get the FP32-FMA roofline denominator (not in MEASURED_PEAKS.json). */
int pvb200_probe_fp32_fma(float* sink, int iters, double* flops_out, pvb200_stream_t stream);
/* the same probe issued as packed fma.rn.f32x2 (SASS FFMA2, two fp32 FMAs per instruction): the convolution kernels use
 * the packed form, so bench.py takes the HIGHER of the two probes as the roofline denominator. */
int pvb200_probe_fp32_fma2(float* sink, int iters, double* flops_out, pvb200_stream_t stream);

/* ---- a1/a2: int16 satellite normalisation ---------------------------------------------------
 * replaces: predict_pv_yield/netcdf_dataset.py:96-101 (astype(float32); - SAT_MEAN; /= SAT_STD)
 *           + the cast at predict_pv_yield/models/conv3d/model.py:113.
 * y[b,c,...] = (float(x[b,c,...]) - mean[c]) / std[c]   two IEEE roundings, true division:
 * bit-identical to the reference fp32 arithmetic.  thw = T*H*W elements per (b,c) plane.
 * _bf16 writes round-to-nearest-even bf16 of that fp32 value. */
int pvb200_sat_normalise_f32(const int16_t* x, float* y, const float* mean, const float* std,
                             int B, int C, long long thw, pvb200_stream_t stream);
int pvb200_sat_normalise_bf16(const int16_t* x, uint16_t* y, const float* mean, const float* std,
                              int B, int C, long long thw, pvb200_stream_t stream);

/* ---- a3/a4: Conv3d 3x3x3, stride 1, padding 0, + bias (+ ReLU) -------------------------------
 * replaces: F.relu(self.sat_conv0(sat_data)) / F.relu(layer(out)), model.py:117-120.
 * x: [B,Cin,Ti,Hi,Wi] fp32, or int16 when x_is_i16 != 0 (normalisation fused into the load, then
 * mean/std must be given);  w: [Cout,Cin,3,3,3];  y: [B,Cout,Ti-2,Hi-2,Wi-2]. */
size_t pvb200_conv3d_workspace_bytes(int Cin, int Cout); /* for _fwd and _dgrad (weight re-layout) */
int pvb200_conv3d_fwd_f32(const void* x, int x_is_i16, const float* mean, const float* std,
                          const float* w, const float* bias, float* y,
                          void* workspace, size_t workspace_bytes,
                          int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu,
                          pvb200_stream_t stream);

/* data gradient (autograd of model.py:118-120): gx = conv_transpose(gz, w), optionally masked by the
 * ReLU of the layer below: gx *= (mask_src > 0).  gz: [B,Cout,Ti-2,Hi-2,Wi-2] (already ReLU-masked
 * gradient w.r.t. the pre-activation), mask_src/gx: [B,Cin,Ti,Hi,Wi] (mask_src may be NULL). */
int pvb200_conv3d_dgrad_f32(const float* gz, const float* w, const float* mask_src, float* gx,
                            void* workspace, size_t workspace_bytes,
                            int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                            pvb200_stream_t stream);

/* ---- bf16 tensor-core path (tcgen05 / TMEM / TMA engine) ----------------------------------------------
 * Activations are "blocked" bf16: [B][Cg][T][H][W][8] with Cg = pvb200_blocked_channel_groups(C) channel groups
 * of 8 (padded with zero channels to an even group count).  Same reference call sites as the _f32 functions
 * (model.py:117-120 and autograd); fp32 master weights are converted per call.
 *   nc_to_blocked_bf16: [B][C][T][H][W] fp32 -> blocked bf16 written into the interior of a tensor padded by
 *                       `pad` on every side of T,H,W (caller zero-fills the border once).
 *   conv3d_fwd_bf16   : yb = relu(conv(xb) + bias), written into the interior of an out_pad-padded blocked tensor.
 *   conv3d_dgrad_bf16 : gx = conv_transpose(gz) * (mask_src > 0); gz_padded is gz zero-padded by 2:
 *                       [B][Cg(Cout)][Ti+2][Hi+2][Wi+2][8]; mask_src (unpadded, [B][Cg(Cin)][Ti][Hi][Wi][8]) may be NULL.
 *                       gx_gzw (may be NULL): a second copy of gx in the weight-gradient operand layout of the layer
 *                       BELOW ([B][Cg(Cin)][Ti][QP][8], pitch Wi + 2, caller zero-fills once), so gx is never re-laid out.
 *   sat_normalise_blocked_bf16: int16 cube -> normalised blocked bf16 (a1 fused with the layout change). */
int pvb200_blocked_channel_groups(int C);
size_t pvb200_conv3d_bf16_workspace_bytes(int Cin, int Cout);
int pvb200_nc_to_blocked_bf16(const float* x, uint16_t* y, int B, int C, int T, int H, int W, int pad,
                              pvb200_stream_t stream);
int pvb200_blocked_to_nc_f32(const uint16_t* x, float* y, int B, int C, int T, int H, int W, pvb200_stream_t stream);
int pvb200_conv3d_fwd_bf16(const uint16_t* xb, const float* w, const float* bias, uint16_t* yb,
                           void* workspace, size_t workspace_bytes,
                           int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int out_pad,
                           pvb200_stream_t stream);
int pvb200_sat_normalise_blocked_bf16(const int16_t* x, uint16_t* yb, const float* mean, const float* std,
                                      int B, int C, int T, int H, int W, pvb200_stream_t stream);
int pvb200_conv3d_dgrad_bf16(const uint16_t* gz_padded, const float* w, const uint16_t* mask_src, uint16_t* gx,
                             uint16_t* gx_gzw, void* workspace, size_t workspace_bytes,
                             int B, int Cin, int Ti, int Hi, int Wi, int Cout, int out_pad,
                             pvb200_stream_t stream);

/* bf16 weight + bias gradient on the tensor cores.  gzw is the pre-activation gradient in blocked bf16 with the INPUT
 * pitch: [B][Cg(Cout)][To][QP][8], position q = ho*Wi + wo, QP = pvb200_conv3d_wgrad_bf16_gz_plane(Hi, Wi) (zero in
 * the wrap columns wo >= Wo and in the tail); pvb200_nc_to_gzw_bf16 builds it from [B][Cout][To][Ho][Wo] fp32. */
long long pvb200_conv3d_wgrad_bf16_gz_plane(int Hi, int Wi);
size_t pvb200_conv3d_wgrad_bf16_workspace_bytes(int Cin, int Cout);
int pvb200_nc_to_gzw_bf16(const float* gz, uint16_t* gzw, int B, int Cout, int To, int Ho, int Wo,
                          pvb200_stream_t stream);
int pvb200_conv3d_wgrad_bf16(const uint16_t* xb, const uint16_t* gzw, float* dw, float* db,
                             void* workspace, size_t workspace_bytes,
                             int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                             pvb200_stream_t stream);

/* weight + bias gradient (autograd of model.py:117-120): dw[co,ci,kt,kh,kw] = sum gz * x(shifted),
 * db[co] = sum gz.  Deterministic two-pass reduction through the caller's workspace. */
size_t pvb200_conv3d_wgrad_workspace_bytes(int Cin, int Cout);
int pvb200_conv3d_wgrad_f32(const void* x, int x_is_i16, const float* mean, const float* std,
                            const float* gz, float* dw, float* db,
                            void* workspace, size_t workspace_bytes,
                            int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                            pvb200_stream_t stream);

/* ---- time-padded variants (conv3d_sat_nwp, nn.Conv3d(..., padding=(1, 0, 0)),
 * predict_pv_yield/models/conv3d/model_sat_nwp.py:85-100,130-143).  pad_t in {0, 1}: To = Ti + 2*pad_t - 2.
 * Ti is always the INPUT (x / gx) time extent; gz has To planes. */
int pvb200_conv3d_fwd_f32_tpad(const void* x, int x_is_i16, const float* mean, const float* std, const float* w,
                               const float* bias, float* y, void* workspace, size_t workspace_bytes,
                               int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int pad_t,
                               pvb200_stream_t stream);
int pvb200_conv3d_dgrad_f32_tpad(const float* gz, const float* w, const float* mask_src, float* gx,
                                 void* workspace, size_t workspace_bytes,
                                 int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t, pvb200_stream_t stream);
int pvb200_conv3d_wgrad_f32_tpad(const void* x, int x_is_i16, const float* mean, const float* std,
                                 const float* gz, float* dw, float* db, void* workspace, size_t workspace_bytes,
                                 int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t, pvb200_stream_t stream);

/* bf16 tensor-core convolutions with time padding pad_t in {0, 1} (the towers of conv3d_sat_nwp in bf16): the planes
 * of the padding are skipped inside the kernel, nothing is padded in memory.  fwd: y has Ti + 2 pad_t - 2 planes.
 * dgrad: gz (Ti + 2 pad_t - 2 planes) arrives zero-padded by 2 on T, H, W as for pvb200_conv3d_dgrad_bf16. */
int pvb200_conv3d_fwd_bf16_tpad(const uint16_t* xb, const float* w, const float* bias, uint16_t* yb, void* workspace,
                                size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu,
                                int out_pad, int pad_t, pvb200_stream_t stream);
int pvb200_conv3d_wgrad_bf16_tpad(const uint16_t* xb, const uint16_t* gzw, float* dw, float* db, void* workspace,
                                  size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                                  pvb200_stream_t stream);
int pvb200_conv3d_dgrad_bf16_tpad(const uint16_t* gz_padded, const float* w, const uint16_t* mask_src, uint16_t* gx,
                                  uint16_t* gx_gzw, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                                  int Hi, int Wi, int Cout, int out_pad, int pad_t, pvb200_stream_t stream);

/* ---- a3/a4/a11 in fp32 MODE on the tensor cores: 3xTF32 implicit GEMM (tcgen05.mma.kind::tf32) -----------------------
 * replaces nn.Conv3d forward / data gradient of model.py:80-90,117-120 at fp32-class accuracy (<= 1e-5): every product is
 * x_hi.w_lo + x_lo.w_hi + x_hi.w_hi with hi = the TF32 truncation the tensor core applies, lo = the exact fp32 residual.
 * Activations are BLOCKED fp32 [B][G][T][H][W][4] (4 channels = 16 bytes innermost, G = pvb200_blocked4_channel_groups(C),
 * even).  Cin, Cout <= 32.  Outputs: a blocked copy (optionally zero-padded by out_pad on T, H, W: the caller keeps the
 * border zero) and / or a plain NCDHW copy [B][Cout][To][Ho][Wo] (what the fp32 head and weight gradients read); either
 * pointer may be null.  pad_t in {0, 1} as for the bf16 kernels.  dgrad: gz arrives blocked and zero-padded by 2 on T, H,
 * W; mask_blk = the blocked activation whose ReLU mask is fused (or null). */
int pvb200_blocked4_channel_groups(int C);
size_t pvb200_conv3d_tf32x3_workspace_bytes(int Cin, int Cout);
/* amax_out (here and below; may be NULL): device scalar that receives max(*amax_out, largest magnitude written), by an
 * atomic max on the bit pattern of the non-negative float -- zero it first.  It is what the two-way fp16 split kernels
 * (pvb200_conv3d_wgrad_f16x2) scale their operands by: producing it in the kernel that writes the tensor saves a pass. */
/* amax_in (convolutions; may be NULL): device scalar max |input tensor| (an amax_out of the kernel that wrote it).  When it
 * is given, Cout > 16 and Cin is in 9..16 or 25..32, the convolution runs the TWO-WAY fp16 split instead of 3xTF32: the input
 * and the weights are scaled by powers of two into fp16's range, every value is split into two fp16 pieces (11 + 11 bits),
 * three of the four piece products run as kind::f16 MMAs over 16 channels at a time -- half the tensor time, the same
 * accuracy class (<= 1e-5) -- and the epilogue scales the sums back. */
int pvb200_nc_to_blocked_f32(const float* x, float* y, int B, int C, int T, int H, int W, int pad, float* amax_out,
                             pvb200_stream_t stream);
int pvb200_blocked_f32_to_nc(const float* x, float* y, int B, int C, int T, int H, int W, pvb200_stream_t stream);
/* a1 fused with the layout change: int16 [B][C][T][H][W] -> normalised blocked fp32 (bit-identical arithmetic) */
int pvb200_sat_normalise_blocked_f32(const int16_t* x, float* y, const float* mean, const float* std, int B, int C, int T,
                                     int H, int W, float* amax_out, pvb200_stream_t stream);
int pvb200_conv3d_fwd_tf32x3(const float* xb, const float* w, const float* bias, float* y_blk, float* y_nc, void* workspace,
                             size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int out_pad,
                             int pad_t, const float* amax_in, float* amax_out, pvb200_stream_t stream);
int pvb200_conv3d_dgrad_tf32x3(const float* gz_padded, const float* w, const float* mask_blk, float* gx_blk, float* gx_nc,
                               void* workspace, size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                               int out_pad, int pad_t, const float* amax_in, float* amax_out, pvb200_stream_t stream);

/* weight / bias gradient of the same layers on the tensor cores (conv3d_wgrad_bf16x3.cu: kind::tf32 cannot read the
 * position-strided operands this reduction needs, so every fp32 value is split EXACTLY into three bf16 pieces and six of
 * the nine piece products run as kind::f16 MMAs -- the same tensor time as 3xTF32, fp32-class accuracy): dw [Cout][Cin][3][3][3],
 * db [Cout] (or null) from x blocked fp32 [B][G(Cin)][Ti][Hi][Wi][4] and the pre-activation gradient gz blocked fp32,
 * zero-padded by gz_pad on T, H, W (2 = the tensor the data gradient reads, 0 = plain).  `supported` says whether the
 * layer fits (Cin, Cout <= 32, rows that fit shared memory); otherwise use pvb200_conv3d_wgrad_f32. */
int pvb200_conv3d_wgrad_bf16x3_supported(int Cin, int Cout, int Hi, int Wi);
size_t pvb200_conv3d_wgrad_bf16x3_workspace_bytes(void);
int pvb200_conv3d_wgrad_bf16x3(const float* xb, const float* gzb, int gz_pad, float* dw, float* db, void* workspace,
                               size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                               pvb200_stream_t stream);
/* the same gradient through a TWO-WAY fp16 split (x s = h0 + h1, 11 + 11 significand bits, three piece products: the
 * accuracy class of the 3xTF32 forward at half the tensor time of the three-way split).  fp16 has 5 exponent bits: the kernel
 * scales each operand by the power of two that brings the tensor's largest magnitude to [2^14, 2^15) and scales the sums
 * back; amax_x / amax_gz = device scalars max |x|, max |gz| (the amax_out of the kernels that wrote the tensors, or
 * pvb200_absmax_f32).  Same shapes, workspace, limits. */
int pvb200_conv3d_wgrad_f16x2(const float* xb, const float* gzb, int gz_pad, const float* amax_x, const float* amax_gz, float* dw, float* db,
                              void* workspace, size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                              int pad_t, pvb200_stream_t stream);
/* *out = max(*out, max_i |x[i]|): out is a device scalar the caller zeroes first */
int pvb200_absmax_f32(const float* x, long long n, float* out, pvb200_stream_t stream);

/* bf16 mode, round 2: weight / bias gradient in the row-step formulation (conv3d_wgrad_bf16_rows.cu): the blocked bf16
 * tensors go from the TMA engine (tensor maps) straight into the operand layout; gz is the tensor the data gradient reads
 * (zero-padded by gz_pad = 2) or a plain one (gz_pad = 0).  `supported`: Cin, Cout <= 32 and rows of at most 64 positions;
 * otherwise use pvb200_conv3d_wgrad_bf16. */
int pvb200_conv3d_wgrad_bf16_rows_supported(int Cin, int Cout, int Hi, int Wi);
size_t pvb200_conv3d_wgrad_bf16_rows_workspace_bytes(void);
int pvb200_conv3d_wgrad_bf16_rows(const uint16_t* xb, const uint16_t* gzb, int gz_pad, float* dw, float* db, void* workspace,
                                  size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                                  pvb200_stream_t stream);

/* ---- general padding (pad_t, pad_hw, pad_hw), each 0 or 1, and MaxPool3d: the Conv3dMaxPool front-end of the Perceiver
 * hybrid (nn.Conv3d(..., padding=(1, 1, 1)) + nn.MaxPool3d(3, stride=(1, 2, 2), padding=(1, 1, 1)),
 * predict_pv_yield/models/perceiver/perceiver_conv3d_nwp_sat.py:42-57).  Ti/Hi/Wi are the INPUT extents. */
int pvb200_conv3d_fwd_f32_pad(const void* x, int x_is_i16, const float* mean, const float* std, const float* w,
                              const float* bias, float* y, void* workspace, size_t workspace_bytes,
                              int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int pad_t, int pad_hw,
                              pvb200_stream_t stream);
int pvb200_conv3d_dgrad_f32_pad(const float* gz, const float* w, const float* mask_src, float* gx,
                                void* workspace, size_t workspace_bytes,
                                int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t, int pad_hw,
                                pvb200_stream_t stream);
int pvb200_conv3d_wgrad_f32_pad(const void* x, int x_is_i16, const float* mean, const float* std,
                                const float* gz, float* dw, float* db, void* workspace, size_t workspace_bytes,
                                int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t, int pad_hw,
                                pvb200_stream_t stream);
/* MaxPool3d(kernel 3, stride (1, 2, 2), padding (1, 1, 1)) over [N = B*C planes][T][H][W] -> [N][T][Ho][Wo],
 * Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1.  fwd also writes the arg-max (flat index t*H*W + h*W + w inside the plane,
 * first maximum in (t, h, w) scan order like torch); bwd routes each gz to its arg-max, deterministically (gather). */
int pvb200_maxpool3d_fwd_f32(const float* x, float* y, int* argmax, long long N, int T, int H, int W, pvb200_stream_t stream);
int pvb200_maxpool3d_bwd_f32(const float* gz, const int* argmax, float* gx, long long N, int T, int H, int W,
                             pvb200_stream_t stream);

/* ---- a6-a9: the fully connected head --------------------------------------------------------
 * replaces model.py:122-154 (reshape, fc1, fc2, PV-history cat, fc_nwp, NWP cat, fc3, fc4) and its
 * autograd.  All matrices fp32, torch Linear layout [out][in]. */
typedef struct pvb200_head {
  size_t struct_size; /* = sizeof(pvb200_head_t), ABI check */
  int B;              /* samples */
  int F1, F2, F3, FO; /* fc1 / fc2 / fc3 out features, forecast_len (fc4 out) */
  int NPV;            /* PV-history features appended to fc2's output (0 = branch off), model.py:130-136 */
  int NNWP;           /* NWP input features (0 = branch off), model.py:139-148 */
  int FNWP;           /* fc_nwp out features (128, model.py:99) */
  int pv_ns;          /* systems per history row: NPV = pv_nt * pv_ns */
  long long K1;       /* cnn_output_size (fc1 in features) */
  long long pv_sb, pv_st; /* PV-history element (b,t,s) lives at pv[b*pv_sb + t*pv_st + s] */
  /* parameters */
  const float *w1, *b1, *w2, *b2, *wn, *bn, *w3, *b3, *w4, *b4;
  /* inputs */
  const float* x;   /* [B,K1]  flattened last conv activation (post-ReLU), NCDHW order (model.py:122) */
  const float* pv;  /* PV / GSP history (may hold NaN: nan_to_num(0), model.py:131) */
  const float* nwp; /* [B,NNWP] */
  /* saved activations (written by fwd, read by bwd) */
  float* h1;  /* [B,F1]  relu(fc1) */
  float* cat; /* [B,F2+NPV+FNWP']  = [relu(fc2) | pv history | relu(fc_nwp)] (FNWP' = FNWP if NNWP else 0) */
  float* h3;  /* [B,F3]  relu(fc3) */
  float* out; /* [B,FO]  forecast */
  /* backward: input gradient and outputs */
  const float* g_out; /* [B,FO] */
  float* g_h3;  /* [B,F3]   grad wrt fc3 pre-activation */
  float* g_cat; /* [B,F2+NPV+FNWP']  grad wrt fc2 / fc_nwp PRE-activations in their slots (pv slot: unused) */
  float* g_h1;  /* [B,F1]   grad wrt fc1 pre-activation */
  float* g_x;   /* [B,K1]   grad wrt the last conv layer's PRE-activation (ReLU mask of x applied) */
  float *dw1, *db1, *dw2, *db2, *dwn, *dbn, *dw3, *db3, *dw4, *db4;
  /* workspace for the split-K fc1 forward */
  void* workspace;
  size_t workspace_bytes;
} pvb200_head_t;

size_t pvb200_head_fwd_workspace_bytes(int B, int F1, long long K1);
/* forward: fills h1, cat, h3, out */
int pvb200_head_fwd_f32(const pvb200_head_t* h, pvb200_stream_t stream);
/* backward: from g_out fills g_h3, g_cat, g_h1, g_x and all dw / db */
int pvb200_head_bwd_f32(const pvb200_head_t* h, pvb200_stream_t stream);

/* tail-only variants for callers that compute fc1 themselves (the bf16 tensor-core path): _tail_fwd consumes S
 * split-K partials [S][B][F1] passed as h->workspace (h->x may be NULL); _tail_bwd produces everything except fc1's
 * weight / data gradient (g_h1 and db1 ARE produced; dw1, x, g_x may be NULL). */
int pvb200_head_tail_fwd_f32(const pvb200_head_t* h, int S, pvb200_stream_t stream);
int pvb200_head_tail_bwd_f32(const pvb200_head_t* h, pvb200_stream_t stream);

/* ---- generic Linear (+ReLU), embedding and history flatten: the heads of conv3d_sat_nwp ------
 * replace self.fc1 .. self.fc4, self.nwp_fc1/2, self.pv_fc1 (nn.Linear), torch.cat, nn.Embedding and
 * .nan_to_num(0).reshape(...) of predict_pv_yield/models/conv3d/model_sat_nwp.py:102-172,196-266.  w is [N][K] (torch
 * layout); x / y / gy / gx rows may be column slices of wider buffers (ld* = row stride in floats), which is how the
 * concatenations are done without copies.  K >= 8192 uses the weight-streaming fc1 kernels and needs contiguous rows. */
size_t pvb200_linear_workspace_bytes(int B, int N, long long K);
int pvb200_linear_fwd_f32(const float* x, long long ldx, const float* w, const float* bias, float* y, long long ldy,
                          int B, long long K, int N, int relu, void* workspace, size_t workspace_bytes,
                          pvb200_stream_t stream);
/* gy: gradient w.r.t. the layer OUTPUT; y: the saved output if the layer has a ReLU (else NULL).  Writes dw, db and,
 * when gx != NULL, gx = g_pre . W -- multiplied by (x > 0) if mask_gx_with_x (x is itself a post-ReLU activation). */
int pvb200_linear_bwd_f32(const float* x, long long ldx, const float* w, const float* y, long long ldy,
                          const float* gy, long long ldgy, float* gx, long long ldgx, int mask_gx_with_x,
                          float* dw, float* db, int B, long long K, int N, void* workspace, size_t workspace_bytes,
                          pvb200_stream_t stream);
/* pieces of a Linear for callers that run the big GEMMs elsewhere (the bf16 tensor-core fc1 of the towers) */
int pvb200_linear_finish_f32(const float* partial, int S, const float* bias, float* y, long long ldy, int B, int N,
                             int relu, pvb200_stream_t stream);
int pvb200_linear_gpre_f32(const float* gy, long long ldgy, const float* y, long long ldy, float* g_pre, float* db,
                           int B, int N, pvb200_stream_t stream);
int pvb200_embedding_fwd_f32(const float* table, const int* ids, float* y, long long ldy, int B, int V, int D,
                             pvb200_stream_t stream);
int pvb200_embedding_bwd_f32(const float* gy, long long ldgy, const int* ids, float* dtable, int B, int V, int D,
                             pvb200_stream_t stream);
int pvb200_history_flatten_f32(const float* src, long long sb, long long st, float* dst, long long lddst, int B, int nt,
                               int ns, pvb200_stream_t stream);

/* ---- a6/a11 in bf16: fc1 as weight-streaming tensor-core GEMMs ------------------------------------------------
 * replaces self.fc1 / F.relu(self.fc1(out)), model.py:92,125 and autograd.  Features are the last conv activation in
 * blocked bf16 [B][Cg][T][H][W][8]; `shadow` is a bf16 copy of fc1.weight permuted to [Cg*T*H*W][128][8]
 * (pvb200_fc1_make_shadow_bf16, once per optimiser step).  B <= 256, fc1_output_features <= 128.
 *   fwd  : partial[s][b][j], s < pvb200_fc1_fwd_bf16_splits()   (feed to pvb200_head_tail_fwd_f32)
 *   dgrad: gradient w.r.t. the last conv layer's pre-activation (ReLU mask fused), written in the two layouts of
 *          pvb200_conv3d_dgrad_bf16 / pvb200_conv3d_wgrad_bf16 (gz_pad zero-bordered by 2, gzw at pitch W+2)
 *   wgrad: dw1 fp32 [F1][K1] in the reference layout */
size_t pvb200_fc1_bf16_shadow_bytes(int Cg, int T, int H, int W);
int pvb200_fc1_make_shadow_bf16(const float* w1, uint16_t* shadow, int F1, int Cg, int T, int H, int W,
                                pvb200_stream_t stream);
/* Adam step on fc1.weight (same arithmetic as pvb200_adam_step_f32) that also rewrites the bf16 shadow in the same pass */
int pvb200_adam_fc1_shadow(float* w1, const float* grad, float* exp_avg, float* exp_avg_sq, uint16_t* shadow,
                           int F1, int Cg, int T, int H, int W,
                           float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                           pvb200_stream_t stream);
/* Data-parallel variant with the optimiser sharded by output feature (rank r owns rows [row_lo, row_lo + nrows) of
 * fc1.weight; the gradient rows come from a reduce-scatter): updates ONLY those rows of w1 / exp_avg / exp_avg_sq and
 * writes their bf16 copy as a contiguous shard [Cg*T*H*W][nrows][8].  After an all-gather of the shards,
 * pvb200_fc1_shadow_from_shards interleaves them into the shadow [Cg*T*H*W][128][8]. */
int pvb200_adam_fc1_shadow_rows(float* w1, const float* grad, float* exp_avg, float* exp_avg_sq, uint16_t* shard,
                                int F1, int Cg, int T, int H, int W, int row_lo, int nrows,
                                float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                                pvb200_stream_t stream);
int pvb200_fc1_shadow_from_shards(const uint16_t* gathered, uint16_t* shadow, int nshards, int nrows, int Cg, int T,
                                  int H, int W, pvb200_stream_t stream);
int pvb200_fc1_fwd_bf16_splits(void);
int pvb200_fc1_fwd_bf16(const uint16_t* xb, const uint16_t* shadow, float* partial, int B, int F1, int Cg, int T, int H,
                        int W, pvb200_stream_t stream);
/* gzw (the gradient in the round-1 weight-gradient kernel's operand layout) may be NULL: the row-step weight gradient reads
 * gz_pad, so the second copy is only written for layers it does not take (planes wider than 64) */
int pvb200_fc1_dgrad_bf16(const float* g1, const uint16_t* shadow, const uint16_t* xb, uint16_t* gz_pad, uint16_t* gzw,
                          int B, int F1, int Cg, int T, int H, int W, pvb200_stream_t stream);
int pvb200_fc1_wgrad_bf16(const float* g1, const uint16_t* xb, float* dw1, int B, int F1, int Cg, int T, int H, int W,
                          pvb200_stream_t stream);

/* ---- a10: loss -------------------------------------------------------------------------------
 * replaces base_model.py:95-103: y = yield[0:B, -FO:, 0] (strided view: element (b,f) at
 * y[b*y_sb + f*y_sf]); losses[0..3] = {nmae (L1, the returned loss), mse, mse_exp, mae_exp} with
 * nowcasting_utils WeightedLosses weights w[FO].  bwd: g[b,f] = gscale * sign(y_hat - y) / (B*FO). */
int pvb200_l1_loss_fwd_f32(const float* y_hat, const float* y, long long y_sb, long long y_sf,
                           const float* weights, float* losses, int B, int FO, pvb200_stream_t stream);
int pvb200_l1_loss_bwd_f32(const float* y_hat, const float* y, long long y_sb, long long y_sf,
                           const float* gscale /* device scalar: upstream grad */, float* g,
                           int B, int FO, pvb200_stream_t stream);

/* ---- validation results on the device (base_model.py:121-136,222-236) ----------------------
 * out [3][B][FO] = {forecast MW = y_hat * capacity, actual MW = y * capacity, capacity}; horizon [2][FO] = per-horizon
 * {mean (y_hat - y)^2, mean |y_hat - y|}.  y / capacity are strided views ([:, -FO:, 0]); capacity may be NULL (= 1).
 * One kernel and one device->host copy replace the reference's four .cpu().numpy() round trips per batch. */
int pvb200_validation_results_f32(const float* y_hat, const float* y, long long y_sb, long long y_sf,
                                  const float* capacity, long long c_sb, long long c_sf, float* out, float* horizon,
                                  int B, int FO, pvb200_stream_t stream);

/* ---- a12: Adam -------------------------------------------------------------------------------
 * replaces torch.optim.Adam(self.parameters(), lr=0.0005).step(), base_model.py:255-257 (single-tensor
 * arithmetic of torch.optim.adam: lerp / addcmul / sqrt / addcdiv, bias-corrected, no weight decay).
 * Multi-tensor: n tensors described by parallel arrays (HOST arrays of device pointers).
 * grad_scale multiplies every gradient first (1/world_size under data parallelism). */
int pvb200_adam_step_f32(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, const long long* numel,
                         float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                         pvb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PVB200_H */
