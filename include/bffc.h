/*
 * bffc.h — C ABI of the B200-native FFT long-convolution engine ("bffc").
 *
 * This is the drop-in boundary for ONE path of HazyResearch/flash-fft-conv: the fused
 * FFT convolution  y = postgate * irfft-like( FFT_N(pad(u*pregate)) * FFT_N(pad(k)) )[:L]
 * behind  FlashFFTConv(seqlen, dtype)(u, k, pregate, postgate).
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repo):
 *   bffc_fwd          <- monarch_conv_forward_* / butterfly_*_forward pybind ops
 *                        (csrc/flashfftconv/monarch.cpp:16-56, called from
 *                        flashfftconv/conv.py:566-1734 and 3239-3853)
 *   bffc_bwd          <- monarch_conv_backward_* ops (monarch.cpp:27-37; conv.py:1737-3233,
 *                        3856-4958) incl. the host-side dk_f.sum(0) and forward recompute
 *                        (monarch_cuda_interface_bwd_bf16.cu:798-808,1107-1114)
 *   bffc_kf_from_filter / bffc_kf_pack* <- torch.fft.fft of the filter + the k_f Monarch digit permutations done in
 *                        Python per call (conv.py:575, :640, :676, :1423-1424, :1632-1633)
 *   bffc_dk_from_dkf / bffc_dkf_unpack* <- their inverses for dk_f + torch.fft.ifft(...).real (conv.py:1817-1820, :1862, :2954)
 *   bffc_plan_*       <- FlashFFTConv.__init__ constant tables (conv.py:72-551)
 *
 * Conventions: plain pointers and sizes only, all data pointers are DEVICE pointers on the
 * current CUDA device, all work is enqueued on the caller's `stream` (the reference used the
 * legacy default stream).  The caller owns every buffer.  Return value 0 = success, non-zero =
 * error; bffc_last_error() gives a message (thread-local).  No CPU fallback exists: on a machine
 * without an sm_100 GPU every compute entry point fails with BFFC_ERR_NO_DEVICE.
 */
#ifndef BFFC_H_
#define BFFC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFFC_ABI_VERSION 3

/* element types of u / y / gates */
#define BFFC_DTYPE_BF16 0
#define BFFC_DTYPE_FP16 1

/* error codes */
#define BFFC_OK 0
#define BFFC_ERR_INVALID 1     /* bad argument (shape, alignment, dtype)            */
#define BFFC_ERR_UNSUPPORTED 2 /* seqlen / option not implemented                    */
#define BFFC_ERR_NO_DEVICE 3   /* no CUDA device, or device is not sm_100            */
#define BFFC_ERR_CUDA 4        /* a CUDA runtime / driver call or a launch failed    */

/* Opaque.  The tables are immutable after creation and bffc_fwd / bffc_bwd / the filter-side entry points may be called
 * concurrently on one plan from several threads and streams.  bffc_fwd_host uses streams, events and a staging order that
 * belong to the plan: calls to it on the same plan must be serialised by the caller. */
typedef struct bffc_plan bffc_plan;

int bffc_abi_version(void);
const char* bffc_last_error(void);

/* 1 if `seqlen` (FFT size N) is supported by this build for `dtype`, else 0. No GPU needed. */
int bffc_supported(int seqlen, int dtype);

/*
 * Create the per-(seqlen, dtype) plan on the current device: DFT matrices, stage twiddles and
 * the k_f layout map live in device memory owned by the plan (replaces the register_buffer()
 * tables of FlashFFTConv.__init__, conv.py:72-551).
 */
int bffc_plan_create(bffc_plan** plan, int seqlen, int dtype);
int bffc_plan_destroy(bffc_plan* plan);
/* FFT size n of the natural-order spectra at this boundary (k_f = rfft(k, n) handed to bffc_kf_pack*, the natural-order
 * dk_f bffc_dkf_unpack* return, the H x n engine buffers): seqlen for seqlen >= 8192; 8192 for the small sizes (256..4096).
 * A small size runs 8192/seqlen batch members per 8192-point unit as independent seqlen-point circular convolutions; its
 * filter spectrum is the 8192-point spectrum of the zero-extended filter sampled at multiples of 8192/seqlen (the pack
 * functions do that), and bffc_dkf_unpack* return the gradient spectrum on the same 8192-point grid (non-zero at those
 * multiples only), so dk = ifft(dk_f, n).real[:, :Lk] holds for every size. */
int bffc_fft_size(const bffc_plan* plan);
/* L passed to bffc_fwd / bffc_bwd / bffc_fwd_host must be a multiple of this: 64 for seqlen <= 8192 (TMA tiles of 64
 * columns), 8 for 16K..512K (16-byte vectors of the CUDA-core outer stage), seqlen/128 for 1M / 2M / 4M (whole rows of the
 * [128][seqlen/128] view of the tcgen05 outer stage).  Other lengths return BFFC_ERR_UNSUPPORTED; a caller holding such a
 * tensor zero-pads it to the next multiple (the operator is unchanged: implicit zero padding), as the host mirror does.
 * The reference itself only requires L even (README.md:270). */
int bffc_length_multiple(const bffc_plan* plan);

/*
 * Frequency-domain filter layout.  The engine consumes k_f = FFT_N(k)/N as packed complex
 * (re, im) pairs in the plan's own digit order ("engine order"), H x N entries, element type
 * = plan dtype (4 bytes per complex entry).
 *
 * bffc_kf_pack:   kf_natural : (H, N) complex64 (interleaved float2), natural frequency order,
 *                 NOT yet scaled.  Writes kf_engine (H*N*4 bytes): engine order, scaled by 1/N,
 *                 optionally conjugated (conj != 0, used by backward for du).
 * bffc_dkf_unpack: dkf_engine : (H, N) float2 in engine order (as written by bffc_bwd)
 *                 -> dkf_natural (H, N) complex64 natural order (conv.py:1818/1862/2954 analogue).
 */
int bffc_kf_pack(const bffc_plan* plan, const void* kf_natural, void* kf_engine, int H, int conj,
                 void* stream);
/* Same as bffc_kf_pack, but kf_half holds only the N/2+1 non-redundant frequencies of the real filter
 * (torch.fft.rfft(k, n=N), complex64); the other half is filled in by Hermitian symmetry. */
int bffc_kf_pack_rfft(const bffc_plan* plan, const void* kf_half, void* kf_engine, int H, int conj,
                      void* stream);
int bffc_dkf_unpack(const bffc_plan* plan, const void* dkf_engine, void* dkf_natural, int H,
                    void* stream);
/* dkf_engine -> dkf_half (H, N/2 + 1) complex64: the non-redundant bins of the Hermitian part (X[k] + conj X[N-k]) / 2 of
 * the gradient spectrum, natural order, so that dk = irfft(dkf_half, n = N)[:, :Lk] (the real part the reference takes of
 * its complex inverse FFT, conv.py:1817-1820, at half the transform work). */
int bffc_dkf_unpack_half(const bffc_plan* plan, const void* dkf_engine, void* dkf_half, int H,
                         void* stream);

/*
 * Filter-side transforms, fp32 on CUDA cores, written / read directly in engine order (no library FFT on the path):
 *   bffc_kf_from_filter: k (H, Lk) fp32 device, Lk <= seqlen  ->  kf_engine  = bffc_kf_pack_rfft(rfft(k, n = fft size))
 *                        (replaces conv.py:572-575 + :640; two real channels share one complex FFT)
 *   bffc_dk_from_dkf:    dkf_engine (H, fft size) float2 as written by bffc_bwd  ->  dk (H, Lk) fp32
 *                        = ifft(unpack(dkf)).real[:, :Lk], small sizes summed over their batch-member blocks (replaces conv.py:1817-1820)
 * Plans with fft size 8192 (seqlen <= 8192): one launch, no workspace (NULL / 0).  Composite sizes N = R * 8192: per group
 * of channels one launch of R-point column FFTs and one of 8192-point row FFTs, with (channels, R/2 + 1, 8192) complex64
 * between them in `workspace`.  bffc_filter_workspace_bytes(plan, H) is the recommended size (a group that stays in L2,
 * at most H channels); any size >= 2 * (R/2 + 1) * 65536 bytes (one channel pair) works, smaller groups = more launches.
 * bffc_last_launch_count() reports the launches of the call.
 */
size_t bffc_filter_workspace_bytes(const bffc_plan* plan, int H);
int bffc_kf_from_filter(const bffc_plan* plan, const void* k, int Lk, void* kf_engine, int H, int conj,
                        void* workspace, size_t workspace_bytes, void* stream);
int bffc_dk_from_dkf(const bffc_plan* plan, const void* dkf_engine, void* dk, int Lk, int H,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Scratch the caller must provide.  bffc_workspace_bytes_ex: exact need of bffc_fwd (backward = 0) or bffc_bwd
 * (backward = 1) for a gated / ungated call; bffc_workspace_bytes: enough for any call with these shapes.
 * seqlen <= 8192: 0, except the gated backward (two (B,H,L) tensors: the gated inputs handed to the dk_f kernel).
 * Composite sizes hold the outer stages' output as 16-bit plane pairs of ceil(B/2)*H*N elements: forward nlev pairs,
 * backward nlev + 1 (nlev = 1 for 16K..64K and 1M, 2 for 128K..512K, 2M, 4M). */
size_t bffc_workspace_bytes(const bffc_plan* plan, int B, int H, int L);
size_t bffc_workspace_bytes_ex(const bffc_plan* plan, int B, int H, int L, int gated, int backward);

/*
 * Forward.  u, y, pregate, postgate: (B, H, L) contiguous, plan dtype, L <= N, L a multiple of
 * bffc_length_multiple().  kf_engine: from bffc_kf_pack* / bffc_kf_from_filter.  pregate/postgate: both NULL or both non-NULL
 * (conv.py:557-558).  y[b,h,:] = postgate * circular_conv_N(pad(u*pregate), pad(k))[:L].
 */
int bffc_fwd(const bffc_plan* plan, const void* u, const void* kf_engine, const void* pregate,
             const void* postgate, void* y, int B, int H, int L, void* workspace,
             size_t workspace_bytes, void* stream);

/*
 * Backward.  dout, u (and gates) as in forward.  kf_engine: the forward's filter spectrum; kf_engine_conj: NULL (the
 * kernels conjugate kf_engine in their pointwise multiply) or a pre-conjugated copy from bffc_kf_pack(..., conj=1), in
 * which case kf_engine may be NULL for an ungated call.
 * Outputs: du (B,H,L) plan dtype; dkf_engine (H, N) float2 fp32 summed over B inside the kernel (overwritten, not
 * accumulated across calls); dpregate/dpostgate (B,H,L) plan dtype when gated (else NULL).  Gated: kf_engine is
 * required (dpostgate = dout * conv(u*pregate, k) is one pass of the forward path; du and dpregate come from one more).
 */
int bffc_bwd(const bffc_plan* plan, const void* dout, const void* u, const void* kf_engine,
             const void* kf_engine_conj, const void* pregate, const void* postgate, void* du,
             void* dkf_engine, void* dpregate, void* dpostgate, int B, int H, int L,
             void* workspace, size_t workspace_bytes, void* stream);

/*
 * Forward on HOST buffers (the reference has no counterpart: its user writes u.cuda() -> conv -> y.cpu(),
 * README.md:108-149, three serial steps on one stream).  u_host, pregate_host, postgate_host, y_host: (B, H, L)
 * contiguous host memory of the plan dtype — page-locked for the copies to overlap; kf_engine: DEVICE, from
 * bffc_kf_pack*.  The batch is cut into chunks of bffc_host_chunk_batch() members (wide rows also over channels,
 * ~12 MB per chunk); chunk c+1 is copied in, chunk c convolved and chunk c-1 copied out at the same time on three
 * internal streams (both PCIe directions busy), all
 * ordered after the work already enqueued on `stream` and joined back into `stream` before the call returns (the
 * call itself is asynchronous like every other entry point).  dev_workspace: device scratch of
 * bffc_host_workspace_bytes() bytes (two staging slots of inputs, output and conv workspace).  The internal streams
 * and events belong to the plan: calls on the same plan must not overlap in time on different caller streams.
 */
int bffc_host_chunk_batch(const bffc_plan* plan, int B, int H, int L);
size_t bffc_host_workspace_bytes(const bffc_plan* plan, int B, int H, int L, int gated);
int bffc_fwd_host(const bffc_plan* plan, const void* u_host, const void* kf_engine,
                  const void* pregate_host, const void* postgate_host, void* y_host, int B, int H,
                  int L, void* dev_workspace, size_t dev_workspace_bytes, void* stream);

/* Number of kernel launches the last bffc_fwd / bffc_bwd / bffc_fwd_host on this thread enqueued (bench.py). */
int bffc_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BFFC_H_ */
