/*
 * C ABI of libwebrtc_ns_b200.so -- the B200 (sm_100a) drop-in for the WebRTC
 * noise-suppression path vendored by templeblock/audioSignalProcess under
 * WebRtc_AMP_Port/webrtc/modules/audio_processing/ns/.
 *
 * Part 1 and 2 are the reference's own entry points, same names, argument
 * meaning and error behaviour; each declaration cites the reference interface
 * it replaces.  Part 3 is the batched entry point the reference does not have
 * (BASELINE.json north_star): N independent streams per call, one CUDA launch.
 * Plain C: opaque handles, plain pointers and sizes; no CUDA or C++ types.
 *
 * Every handle owns a slot in a per-GPU structure-of-arrays state slab; the
 * single-stream calls are batches of one (correct, slow), the batch calls are
 * the product.  There is no CPU implementation behind these symbols: if CUDA or
 * a GPU is missing the calls fail (-1) and WebRtcNsB200_LastError() says why.
 */
#ifndef WEBRTC_NS_B200_H_
#define WEBRTC_NS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 1. float suppressor: ns/include/noise_suppression.h ------------------ */
typedef struct NsHandleT NsHandle;

/* noise_suppression.h:35  (ns/noise_suppression.c:20).  0 ok, -1 error. */
int WebRtcNs_Create(NsHandle** NS_inst);
/* noise_suppression.h:48  (noise_suppression.c:31).  Always 0. */
int WebRtcNs_Free(NsHandle* NS_inst);
/* noise_suppression.h:65  (ns_core.c:74): fs in {8000,16000,32000,48000};
 * resets all state and selects policy 0.  0 ok, -1 on NULL / bad fs. */
int WebRtcNs_Init(NsHandle* NS_inst, uint32_t fs);
/* noise_suppression.h:80  (ns_core.c:1013): mode 0..3.  0 ok, -1 otherwise. */
int WebRtcNs_set_policy(NsHandle* NS_inst, int mode);
/* noise_suppression.h:93  (ns_core.c:1043): 10 ms of band 0 (80 samples at
 * 8 kHz, else 160), float in int16 scale.  The statistics update is fused into
 * WebRtcNs_Process here (all reference callers pass the same frame to both:
 * test_ns_module.cpp:97-99, libapm/src/apm_ns.cpp:69-74); this call only
 * records the frame so that Process can tell whether the caller changed it. */
void WebRtcNs_Analyze(NsHandle* NS_inst, const float* spframe);
/* noise_suppression.h:108 (ns_core.c:1183): num_bands in 1..3 pointers of one
 * 10 ms frame each; outframe may alias spframe.  Output clamped to
 * [-32768, 32767]. */
void WebRtcNs_Process(NsHandle* NS_inst, const float* const* spframe, int num_bands,
                      float* const* outframe);
/* noise_suppression.h:123 (noise_suppression.c:57): -1 on NULL / uninitialised. */
float WebRtcNs_prior_speech_probability(NsHandle* handle);

/* ---- 2. fixed-point suppressor: ns/include/noise_suppression_x.h ---------- */
typedef struct NsxHandleT NsxHandle;

/* noise_suppression_x.h (ns/noise_suppression_x.c:19-54, nsx_core.c:630,785,1502) */
int WebRtcNsx_Create(NsxHandle** nsxInst);
int WebRtcNsx_Free(NsxHandle* nsxInst);
int WebRtcNsx_Init(NsxHandle* nsxInst, uint32_t fs);
int WebRtcNsx_set_policy(NsxHandle* nsxInst, int mode);
void WebRtcNsx_Process(NsxHandle* nsxInst, const short* const* speechFrame, int num_bands,
                       short* const* outFrame);

/* ---- 3. batched entry points (new) ---------------------------------------- */
/*
 * For every stream s < n_streams and frame f < frames: Analyze + Process of the
 * frame at pcm_in[s*in_stride + f*(fs/100)] into pcm_out[s*out_stride +
 * f*(fs/100)], exactly what the reference callers do per frame
 * (apm_ns.cpp:96-132): full-band int16 PCM in, int16 out; at 32/48 kHz the
 * band split/merge of AudioBuffer (audio_buffer.cc:455-463, splitting_filter.cc)
 * runs on the GPU around the suppressor.  All handles must share one sample
 * rate; they may live on different GPUs (the call buckets them by device).
 * Strides are in samples and must be even.  pcm_out may alias pcm_in.
 * Host-pointer version: copies in, runs, copies out, returns when done.
 * 0 ok, -1 error (see WebRtcNsB200_LastError).
 */
int WebRtcNs_ProcessBatch(NsHandle* const* handles, int n_streams, const int16_t* pcm_in,
                          size_t in_stride, int16_t* pcm_out, size_t out_stride, int frames);
int WebRtcNsx_ProcessBatch(NsxHandle* const* handles, int n_streams, const int16_t* pcm_in,
                           size_t in_stride, int16_t* pcm_out, size_t out_stride, int frames);
/* Asynchronous form of WebRtcNs[x]_ProcessBatch for callers that stream: same arguments and
 * results, but the call returns as soon as its copies and kernels are enqueued, and *ticket names
 * it.  `in` must stay unchanged and `out` unread until WebRtcNsB200_WaitBatch(ticket) returns.
 * Consecutive asynchronous calls form ONE pipeline over the GPU's two copy engines: the copy-out
 * of call k overlaps the copy-in and kernels of call k+1, so a caller that keeps two calls in
 * flight (two buffer pairs) is bound by PCIe duplex bandwidth, not by each call's fill and drain.
 * Host buffers must be page-locked (cudaHostAlloc / cudaHostRegister) -- from pageable memory
 * CUDA copies synchronously and the call degenerates to the blocking one.  Any other library
 * call first waits for every batch still in flight, so mixing the two forms is safe.  At 32/48
 * kHz the call blocks (ticket 0).  WaitBatch(0) and waiting twice are no-ops. */
int WebRtcNs_ProcessBatchAsync(NsHandle* const* handles, int n_streams, const int16_t* pcm_in,
                               size_t in_stride, int16_t* pcm_out, size_t out_stride, int frames,
                               uint64_t* ticket);
int WebRtcNsx_ProcessBatchAsync(NsxHandle* const* handles, int n_streams, const int16_t* pcm_in,
                                size_t in_stride, int16_t* pcm_out, size_t out_stride, int frames,
                                uint64_t* ticket);
int WebRtcNsB200_WaitBatch(uint64_t ticket);

/*
 * Device-pointer version: pcm_in/pcm_out are device memory on the GPU that owns
 * ALL the handles; the work is enqueued on cuda_stream (a cudaStream_t cast to
 * void*, NULL = the library's own stream) and the call returns without
 * synchronising.
 */
int WebRtcNs_ProcessBatchDevice(NsHandle* const* handles, int n_streams, const int16_t* pcm_in,
                                size_t in_stride, int16_t* pcm_out, size_t out_stride, int frames,
                                void* cuda_stream);
int WebRtcNsx_ProcessBatchDevice(NsxHandle* const* handles, int n_streams, const int16_t* pcm_in,
                                 size_t in_stride, int16_t* pcm_out, size_t out_stride,
                                 int frames, void* cuda_stream);
/*
 * Float mirror of WebRtcNs_Process over a batch: band frames in int16-scale
 * float, layout [stream][frame][band][frame_len] with the given strides (in
 * floats); host pointers.
 */
int WebRtcNs_ProcessBatchBandsF32(NsHandle* const* handles, int n_streams, int num_bands,
                                  const float* in, size_t in_stream_stride, float* out,
                                  size_t out_stream_stride, int frames);
/*
 * Analyze and Process fed different signals (float suppressor): full APM places the echo
 * canceller between WebRtcNs_Analyze and WebRtcNs_Process (audio_processing_impl.cc:625-631), so
 * the noise statistics come from `analyze_in` while `pcm_in` is what gets filtered.  The
 * single-stream WebRtcNs_Analyze / WebRtcNs_Process pair does the same when their frames differ.
 * Full-band int16 PCM versions: 8 and 16 kHz (one band); all handles on one GPU.  Band-frame
 * float version: any rate, analyze_in holds band-0 frames [stream][frame][frame_len].
 * A stream that has been driven this way stays on the two-signal kernel (later fused calls feed
 * the Process signal to both sides, with results identical to the reference's).
 */
int WebRtcNs_AnalyzeProcessBatch(NsHandle* const* handles, int n_streams, const int16_t* analyze_in,
                                 size_t analyze_stride, const int16_t* pcm_in, size_t in_stride,
                                 int16_t* pcm_out, size_t out_stride, int frames);
int WebRtcNs_AnalyzeProcessBatchDevice(NsHandle* const* handles, int n_streams,
                                       const int16_t* analyze_in, size_t analyze_stride,
                                       const int16_t* pcm_in, size_t in_stride, int16_t* pcm_out,
                                       size_t out_stride, int frames, void* cuda_stream);
int WebRtcNs_AnalyzeProcessBatchBandsF32(NsHandle* const* handles, int n_streams, int num_bands,
                                         const float* analyze_in, size_t analyze_stream_stride,
                                         const float* in, size_t in_stream_stride, float* out,
                                         size_t out_stream_stride, int frames);
/*
 * Interleaved multi-channel capture block, processed in place: the job of the author's wrapper
 * APM_NS::processCaptureStream (libapm/src/apm_ns.cpp:91-132 short, :47-89 float) -- deinterleave
 * (float samples in [-1, 1] through FloatToS16, audio_util.h:27-32), band split, Analyze + Process
 * per channel (handles[c] = channel c), merge, interleave (S16ToFloat) -- with the (de)interleave
 * and conversions done on the GPU.  samples_per_channel = frames * fs/100, a multiple of 8.
 */
int WebRtcNs_ProcessInterleavedI16(NsHandle* const* handles, int n_channels, int16_t* data,
                                   int samples_per_channel);
int WebRtcNs_ProcessInterleavedF32(NsHandle* const* handles, int n_channels, float* data,
                                   int samples_per_channel);
/* Init + set_policy for many handles with one launch (same effect as calling
 * WebRtcNs_Init / WebRtcNs_set_policy on each). */
int WebRtcNs_InitBatch(NsHandle* const* handles, int n_streams, uint32_t fs, int mode);
int WebRtcNsx_InitBatch(NsxHandle* const* handles, int n_streams, uint32_t fs, int mode);

/* ---- 4. library utilities -------------------------------------------------- */
/* GPU on which subsequent Create calls place their stream (default: the
 * calling thread's current CUDA device).  -1 = back to default. */
int WebRtcNsB200_SetCreateDevice(int device);
int WebRtcNsB200_DeviceCount(void);
/* Waits for everything the library enqueued on every device. */
int WebRtcNsB200_Synchronize(void);
const char* WebRtcNsB200_LastError(void);
/* Kernels launched by this library since load (for bench.py's gpu_launches). */
uint64_t WebRtcNsB200_KernelLaunches(void);
/* Device self-test of the kernels' arithmetic shortcuts against their exact definitions over
 * n_cases operands: logf, sqrtf, int16 rounding, the floor square root and the packed (f32x2) forms
 * of multiply-add, complex multiply, division and square root must be identical to the scalar ones;
 * the branch-free division must equal IEEE division except for at most 2 per million quotients
 * that may be one ulp off; the tracker's logarithm must be the double-precision logarithm rounded
 * to float, likewise the exponential and the sigmoid map built on tanh (never more than an ulp off, at
 * most 1 in 100 000 not identical).  stats[8]: [0] hard mismatches, [1] divisions one ulp off,
 * [2] divisions checked, [3] logarithms not identical, [4] logarithms checked, [5] exponentials not
 * identical, [6] exponentials checked, [7] sigmoid maps not identical (of [6]).  0 = pass. */
int WebRtcNsB200_SelfTest(uint64_t n_cases);
int WebRtcNsB200_SelfTestStats(uint64_t n_cases, uint64_t* stats);
/* Deterministic synthetic PCM (csrc/pcm_synth.h) written on the device:
 * stream s, sample n -> dst[s*stride + n], n < n_samples, as stream index
 * first_stream + s at time offset first_sample. */
int WebRtcNsB200_SynthPcmDevice(int16_t* dst, size_t stride, int n_streams, uint32_t first_stream,
                                uint32_t fs, uint32_t first_sample, uint32_t n_samples,
                                uint32_t base_seed, void* cuda_stream);
/* Same generator on the host (for baselines and tests). */
void WebRtcNsB200_SynthPcmHost(int16_t* dst, uint32_t stream, uint32_t fs, uint32_t first_sample,
                               uint32_t n_samples, uint32_t base_seed);
/* Per-stream checksum (sum of samples, sum of squares) of device PCM, for
 * workloads too large to read back.
 * sums: [n_streams][2] int64 on the device. */
int WebRtcNsB200_ChecksumDevice(const int16_t* pcm, size_t stride, int n_streams,
                                uint32_t n_samples, int64_t* sums, void* cuda_stream);
/* Same, added onto the sums already there: a job cut into chunks (config 5 generates, processes
 * and reduces 10 minutes of PCM chunk by chunk) keeps one running checksum per stream. */
int WebRtcNsB200_ChecksumAccumulateDevice(const int16_t* pcm, size_t stride, int n_streams,
                                          uint32_t n_samples, int64_t* sums, void* cuda_stream);

/* ---- 5. stream state snapshot / restore / migration --------------------------- */
/* No reference counterpart (the reference's state is a malloc'ed struct the caller could memcpy:
 * NoiseSuppressionC, ns_core.h:52-114; NoiseSuppressionFixedC, nsx_core.h:22-110); here the state
 * lives in GPU slabs, so long-lived streams get explicit calls.  `handle` is an NsHandle* or an
 * NsxHandle*.  Export waits for the stream's enqueued work; Import restores bit-exactly into any
 * handle of the same kind (on any GPU), after which processing continues as if uninterrupted.
 * Migrate moves the stream's slabs to another GPU (peer copy) and keeps the handle valid. */
size_t WebRtcNsB200_StateSize(const void* handle);              /* bytes; 0 on a bad handle */
int WebRtcNsB200_ExportState(const void* handle, void* buf, size_t size);
int WebRtcNsB200_ImportState(void* handle, const void* buf, size_t size);
int WebRtcNsB200_MigrateHandle(void* handle, int device);
int WebRtcNsB200_HandleDevice(const void* handle);              /* -1 on a bad handle */

#ifdef __cplusplus
}
#endif

#endif /* WEBRTC_NS_B200_H_ */
