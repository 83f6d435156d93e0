// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin driver over the UNMODIFIED reference sources (compiled where they lie
// under /root/reference by oracle/Makefile into oracle/_ref/libns_ref.so).
// Nothing in this file restates the algorithm: it only walks 10 ms frames
// through the reference's own public API, exactly as its three call sites do
// (WebRtc_AMP_Port/test_ns_module.cpp:83-109,
//  WebRtc_AMP_Port/libapm/src/apm_ns.cpp:96-132,
//  webrtc/modules/audio_processing/noise_suppression_impl.cc:66-95).
//
// Entry points (all extern "C", plain pointers):
//   ref_ns_run      float NS  (WebRtcNs_Analyze + WebRtcNs_Process), 8/16/32/48 kHz
//   ref_nsx_run     fixed NSx (WebRtcNsx_Process),                   8/16/32/48 kHz
//   ref_ns_run_mt / ref_nsx_run_mt   same over many streams with pthreads
//                   (the CPU baseline of bench.py)
//   ref_batch_*     persistent handles + a thread pool: the CPU arm of bench.py (handles live across
//                   steps like the GPU arm's, streams handed to threads dynamically)
//   ref_ns_trace    float NS with the per-frame decision state copied out of the reference struct
//   ref_rdft        WebRtc_rdft itself (the oracle's FFT restatement is pinned bit for bit against it)
//   ref_synth_pcm   csrc/pcm_synth.h on the host, so that the reference arm never loads the product library
//   ref_qmf_* / ref_resample_* / ref_spl_*  primitive hooks for the KATs.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <vector>

#include "webrtc/common_audio/resampler/push_sinc_resampler.h"
#include "webrtc/common_audio/signal_processing/include/real_fft.h"
#include "webrtc/common_audio/signal_processing/include/signal_processing_library.h"
#include "libapm/include/apm_ns.h"
#include "webrtc/modules/audio_processing/audio_buffer.h"
#include "webrtc/modules/audio_processing/ns/include/noise_suppression.h"
#include "webrtc/modules/audio_processing/ns/include/noise_suppression_x.h"
extern "C" {
#include "webrtc/modules/audio_processing/ns/ns_core.h"        // state trace (read-only) for the parity triage
#include "webrtc/modules/audio_processing/utility/fft4g.h"     // WebRtc_rdft hook: pins the oracle's FFT restatement
}
#include "../audiosignalprocess_b200/csrc/pcm_synth.h"         // our synthetic-PCM generator (integer only, header only)

using webrtc::AudioBuffer;

namespace {

inline int FrameLen(int fs) { return fs / 100; }

inline int16_t RoundToS16(float v) {  // audio_util.h:41-49 contract, via the reference's own helper
  return webrtc::FloatS16ToS16(v);
}

}  // namespace

extern "C" {

// Float NS over one stream. pcm_in: nframes*fs/100 int16. Outputs (any may be
// NULL):
//   out_f32    8/16 kHz only: the raw float frame WebRtcNs_Process wrote
//              (int16-scale floats, clamped, not rounded). Not written at
//              32/48 kHz, where AudioBuffer merges the bands in int16.
//   out_i16    the int16 frame a caller of AudioBuffer reads back.
//   prior_prob WebRtcNs_prior_speech_probability after each frame.
int ref_ns_run(int fs, int mode, int nframes, const int16_t* pcm_in,
               float* out_f32, int16_t* out_i16, float* prior_prob) {
  NsHandle* h = NULL;
  if (WebRtcNs_Create(&h) != 0) return -1;
  if (WebRtcNs_Init(h, (uint32_t)fs) != 0 || WebRtcNs_set_policy(h, mode) != 0) {
    WebRtcNs_Free(h);
    return -1;
  }
  const int n = FrameLen(fs);
  if (fs == 8000 || fs == 16000) {
    std::vector<float> in(n), out(n);
    for (int f = 0; f < nframes; ++f) {
      for (int i = 0; i < n; ++i) in[i] = (float)pcm_in[(size_t)f * n + i];
      const float* inb[1] = {in.data()};
      float* outb[1] = {out.data()};
      WebRtcNs_Analyze(h, in.data());
      WebRtcNs_Process(h, inb, 1, outb);
      for (int i = 0; i < n; ++i) {
        if (out_f32) out_f32[(size_t)f * n + i] = out[i];
        if (out_i16) out_i16[(size_t)f * n + i] = RoundToS16(out[i]);
      }
      if (prior_prob) prior_prob[f] = WebRtcNs_prior_speech_probability(h);
    }
  } else {
    AudioBuffer ab(n, 1, n, 1, n);
    for (int f = 0; f < nframes; ++f) {
      memcpy(ab.data(0), pcm_in + (size_t)f * n, sizeof(int16_t) * n);
      ab.SplitIntoFrequencyBands();
      WebRtcNs_Analyze(h, ab.split_bands_const_f(0)[webrtc::kBand0To8kHz]);
      WebRtcNs_Process(h, ab.split_bands_const_f(0), ab.num_bands(), ab.split_bands_f(0));
      ab.MergeFrequencyBands();
      if (out_i16) memcpy(out_i16 + (size_t)f * n, ab.data_const(0), sizeof(int16_t) * n);
      if (prior_prob) prior_prob[f] = WebRtcNs_prior_speech_probability(h);
    }
  }
  WebRtcNs_Free(h);
  return 0;
}

// Float NS with different signals for Analyze and Process (an echo canceller between them:
// audio_processing_impl.cc:625-631), straight through the reference API.  ana: [frame][fl] band-0
// frames for WebRtcNs_Analyze; in/out: [frame][band][fl] band frames for WebRtcNs_Process
// (int16-scale floats).  The first fused_frames frames feed band 0 of `in` to both.
int ref_ns_split_run(int fs, int mode, int nb, int nframes, int fused_frames, const float* ana,
                     const float* in, float* out) {
  NsHandle* h = NULL;
  if (WebRtcNs_Create(&h) != 0) return -1;
  if (WebRtcNs_Init(h, (uint32_t)fs) != 0 || WebRtcNs_set_policy(h, mode) != 0) {
    WebRtcNs_Free(h);
    return -1;
  }
  const int fl = fs == 8000 ? 80 : 160;
  for (int f = 0; f < nframes; ++f) {
    const float* bands_in[3];
    float* bands_out[3];
    for (int b = 0; b < nb; ++b) {
      bands_in[b] = in + ((size_t)f * nb + b) * fl;
      bands_out[b] = out + ((size_t)f * nb + b) * fl;
    }
    WebRtcNs_Analyze(h, f < fused_frames ? bands_in[0] : ana + (size_t)f * fl);
    WebRtcNs_Process(h, bands_in, nb, bands_out);
  }
  WebRtcNs_Free(h);
  return 0;
}

// Fixed NSx over one stream (explicit band split at 32/48 kHz, as
// noise_suppression_impl.cc:90-93 would do; the author's -DNS_FIXED driver
// never splits -- SURVEY.md appendix B.9).
int ref_nsx_run(int fs, int mode, int nframes, const int16_t* pcm_in, int16_t* out_i16) {
  NsxHandle* h = NULL;
  if (WebRtcNsx_Create(&h) != 0) return -1;
  if (WebRtcNsx_Init(h, (uint32_t)fs) != 0 || WebRtcNsx_set_policy(h, mode) != 0) {
    WebRtcNsx_Free(h);
    return -1;
  }
  const int n = FrameLen(fs);
  if (fs == 8000 || fs == 16000) {
    std::vector<int16_t> in(n), out(n);
    for (int f = 0; f < nframes; ++f) {
      memcpy(in.data(), pcm_in + (size_t)f * n, sizeof(int16_t) * n);
      const short* inb[1] = {in.data()};
      short* outb[1] = {out.data()};
      WebRtcNsx_Process(h, inb, 1, outb);
      memcpy(out_i16 + (size_t)f * n, out.data(), sizeof(int16_t) * n);
    }
  } else {
    AudioBuffer ab(n, 1, n, 1, n);
    for (int f = 0; f < nframes; ++f) {
      memcpy(ab.data(0), pcm_in + (size_t)f * n, sizeof(int16_t) * n);
      ab.SplitIntoFrequencyBands();
      WebRtcNsx_Process(h, ab.split_bands_const(0), ab.num_bands(), ab.split_bands(0));
      ab.MergeFrequencyBands();
      memcpy(out_i16 + (size_t)f * n, ab.data_const(0), sizeof(int16_t) * n);
    }
  }
  WebRtcNsx_Free(h);
  return 0;
}

// ---- many streams, pthreads: one stream per core at a time ------------------
struct MtJob {
  int fixed, fs, mode, nframes, nstreams, tid, nthreads;
  const int16_t* in;
  int16_t* out;
};

static void* MtWorker(void* p) {
  MtJob* j = (MtJob*)p;
  const size_t per = (size_t)j->nframes * FrameLen(j->fs);
  for (int s = j->tid; s < j->nstreams; s += j->nthreads) {
    if (j->fixed)
      ref_nsx_run(j->fs, j->mode, j->nframes, j->in + s * per, j->out + s * per);
    else
      ref_ns_run(j->fs, j->mode, j->nframes, j->in + s * per, NULL, j->out + s * per, NULL);
  }
  return NULL;
}

// pcm_in/out: [nstreams][nframes*fs/100] int16. Returns wall seconds spent in
// the processing loop only (CLOCK_MONOTONIC), or <0 on error.
double ref_run_mt(int fixed, int fs, int mode, int nstreams, int nframes, int nthreads,
                  const int16_t* pcm_in, int16_t* pcm_out) {
  if (nthreads < 1) nthreads = 1;
  std::vector<pthread_t> th(nthreads);
  std::vector<MtJob> jobs(nthreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; ++t) {
    MtJob j = {fixed, fs, mode, nframes, nstreams, t, nthreads, pcm_in, pcm_out};
    jobs[t] = j;
    if (pthread_create(&th[t], NULL, MtWorker, &jobs[t]) != 0) return -1.0;
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// ---- persistent batch: the CPU arm of bench.py --------------------------------------------------
// One reference handle (and, at 32/48 kHz, one AudioBuffer) per stream, created once and kept
// across steps exactly like the GPU arm keeps its streams (test_ns_module.cpp:62-109: one handle,
// one loop).  A pool of threads started once takes streams one at a time from a shared counter
// ("one stream per core at a time"), so a silent stream class cannot idle a thread while another
// thread still holds a queue of expensive ones.
struct RefBatch {
  int fixed, fs, mode, n, nthreads;
  std::vector<void*> handles;
  std::vector<AudioBuffer*> bufs;     // 32/48 kHz only
  std::vector<pthread_t> threads;
  pthread_mutex_t mu;
  pthread_cond_t cv_go, cv_done;
  unsigned long long generation;
  int running, quit;
  // the step in flight
  int frames;
  const int16_t* in;
  int16_t* out;
  size_t in_stride, out_stride;
  volatile int next;
};

static void RefBatchStream(RefBatch* b, int s) {
  const int n = FrameLen(b->fs);
  const int16_t* in = b->in + (size_t)s * b->in_stride;
  int16_t* out = b->out + (size_t)s * b->out_stride;
  if (b->fs <= 16000) {
    if (b->fixed) {
      NsxHandle* h = (NsxHandle*)b->handles[s];
      for (int f = 0; f < b->frames; ++f) {
        const short* inb[1] = {in + (size_t)f * n};
        short* outb[1] = {out + (size_t)f * n};
        WebRtcNsx_Process(h, inb, 1, outb);
      }
    } else {
      NsHandle* h = (NsHandle*)b->handles[s];
      float fin[160], fout[160];
      for (int f = 0; f < b->frames; ++f) {
        for (int i = 0; i < n; ++i) fin[i] = (float)in[(size_t)f * n + i];
        const float* inb[1] = {fin};
        float* outb[1] = {fout};
        WebRtcNs_Analyze(h, fin);
        WebRtcNs_Process(h, inb, 1, outb);
        for (int i = 0; i < n; ++i) out[(size_t)f * n + i] = RoundToS16(fout[i]);
      }
    }
    return;
  }
  AudioBuffer& ab = *b->bufs[s];
  for (int f = 0; f < b->frames; ++f) {
    memcpy(ab.data(0), in + (size_t)f * n, sizeof(int16_t) * n);
    ab.SplitIntoFrequencyBands();
    if (b->fixed) {
      WebRtcNsx_Process((NsxHandle*)b->handles[s], ab.split_bands_const(0), ab.num_bands(), ab.split_bands(0));
    } else {
      NsHandle* h = (NsHandle*)b->handles[s];
      WebRtcNs_Analyze(h, ab.split_bands_const_f(0)[webrtc::kBand0To8kHz]);
      WebRtcNs_Process(h, ab.split_bands_const_f(0), ab.num_bands(), ab.split_bands_f(0));
    }
    ab.MergeFrequencyBands();
    memcpy(out + (size_t)f * n, ab.data_const(0), sizeof(int16_t) * n);
  }
}

static void* RefBatchWorker(void* p) {
  RefBatch* b = (RefBatch*)p;
  unsigned long long seen = 0;
  for (;;) {
    pthread_mutex_lock(&b->mu);
    while (!b->quit && b->generation == seen) pthread_cond_wait(&b->cv_go, &b->mu);
    if (b->quit) { pthread_mutex_unlock(&b->mu); return NULL; }
    seen = b->generation;
    pthread_mutex_unlock(&b->mu);
    for (;;) {
      const int s = __sync_fetch_and_add(&b->next, 1);
      if (s >= b->n) break;
      RefBatchStream(b, s);
    }
    pthread_mutex_lock(&b->mu);
    if (--b->running == 0) pthread_cond_signal(&b->cv_done);
    pthread_mutex_unlock(&b->mu);
  }
}

void ref_batch_free(void* p);

void* ref_batch_create(int fixed, int fs, int mode, int nstreams, int nthreads) {
  if (nstreams <= 0) return NULL;
  if (nthreads < 1) nthreads = 1;
  RefBatch* b = new RefBatch();
  b->fixed = fixed; b->fs = fs; b->mode = mode; b->n = nstreams; b->nthreads = nthreads;
  b->generation = 0; b->running = 0; b->quit = 0; b->next = 0;
  pthread_mutex_init(&b->mu, NULL);
  pthread_cond_init(&b->cv_go, NULL);
  pthread_cond_init(&b->cv_done, NULL);
  b->handles.assign(nstreams, (void*)NULL);
  bool ok = true;
  for (int s = 0; s < nstreams && ok; ++s) {
    if (fixed) {
      NsxHandle* h = NULL;
      ok = WebRtcNsx_Create(&h) == 0 && WebRtcNsx_Init(h, (uint32_t)fs) == 0 && WebRtcNsx_set_policy(h, mode) == 0;
      b->handles[s] = h;
    } else {
      NsHandle* h = NULL;
      ok = WebRtcNs_Create(&h) == 0 && WebRtcNs_Init(h, (uint32_t)fs) == 0 && WebRtcNs_set_policy(h, mode) == 0;
      b->handles[s] = h;
    }
    if (fs > 16000) b->bufs.push_back(new AudioBuffer(FrameLen(fs), 1, FrameLen(fs), 1, FrameLen(fs)));
  }
  if (!ok) { ref_batch_free(b); return NULL; }
  b->threads.resize(nthreads);
  for (int t = 0; t < nthreads; ++t)
    if (pthread_create(&b->threads[t], NULL, RefBatchWorker, b) != 0) { b->threads.resize(t); ref_batch_free(b); return NULL; }
  return b;
}

// One step: `frames` 10 ms frames of every stream.  in/out: [stream][stride] int16 (out may alias in).
// Returns the wall seconds of the step (CLOCK_MONOTONIC around wake-up .. last thread done).
double ref_batch_step(void* p, int frames, const int16_t* in, size_t in_stride, int16_t* out, size_t out_stride) {
  RefBatch* b = (RefBatch*)p;
  if (!b || frames < 0) return -1.0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_mutex_lock(&b->mu);
  b->frames = frames; b->in = in; b->out = out; b->in_stride = in_stride; b->out_stride = out_stride;
  b->next = 0;
  b->running = (int)b->threads.size();
  b->generation++;
  pthread_cond_broadcast(&b->cv_go);
  while (b->running > 0) pthread_cond_wait(&b->cv_done, &b->mu);
  pthread_mutex_unlock(&b->mu);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_batch_free(void* p) {
  RefBatch* b = (RefBatch*)p;
  if (!b) return;
  pthread_mutex_lock(&b->mu);
  b->quit = 1;
  pthread_cond_broadcast(&b->cv_go);
  pthread_mutex_unlock(&b->mu);
  for (size_t t = 0; t < b->threads.size(); ++t) pthread_join(b->threads[t], NULL);
  for (size_t s = 0; s < b->handles.size(); ++s) {
    if (!b->handles[s]) continue;
    if (b->fixed) WebRtcNsx_Free((NsxHandle*)b->handles[s]);
    else WebRtcNs_Free((NsHandle*)b->handles[s]);
  }
  for (size_t s = 0; s < b->bufs.size(); ++s) delete b->bufs[s];
  pthread_mutex_destroy(&b->mu);
  pthread_cond_destroy(&b->cv_go);
  pthread_cond_destroy(&b->cv_done);
  delete b;
}

// Synthetic PCM of streams [first_stream, first_stream + nstreams) on `nthreads` host threads
// (csrc/pcm_synth.h is a pure function of (seed, stream, fs, n): the same bits as the device generator).
struct SynthJob { int16_t* dst; size_t stride; int n, tid, nthreads; uint32_t first_stream, fs, first_sample, n_samples, seed; };
static void* SynthWorker(void* p) {
  SynthJob* j = (SynthJob*)p;
  for (int s = j->tid; s < j->n; s += j->nthreads) {
    int16_t* d = j->dst + (size_t)s * j->stride;
    for (uint32_t i = 0; i < j->n_samples; ++i) d[i] = pcm_synth_sample(j->seed, j->first_stream + (uint32_t)s, j->fs, j->first_sample + i);
  }
  return NULL;
}
void ref_synth_pcm(int16_t* dst, size_t stride, int nstreams, uint32_t first_stream, uint32_t fs, uint32_t first_sample,
                   uint32_t n_samples, uint32_t seed, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<pthread_t> th(nthreads);
  std::vector<SynthJob> jobs(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    SynthJob j = {dst, stride, nstreams, t, nthreads, first_stream, fs, first_sample, n_samples, seed};
    jobs[t] = j;
    pthread_create(&th[t], NULL, SynthWorker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

// ---- float NS with the decision state copied out of the reference's struct after every frame ---
// (8/16 kHz).  Per frame kRefTraceWords floats, bins padded to 129:
//   [0,387) lquantile[3][129] | [387,774) density[3][129] | quantile | smooth | noisePrev | magnPrevAnalyze |
//   logLrtTimeAvg | magnAvgPause (6 x 129 from 774) | featureData[7] | priorModelPars[7] | priorSpeechProb |
//   blockInd | counter[3] | updates | modelUpdatePars[4] | signalEnergy | sumMagn
int ref_ns_trace_words(void) { return 774 + 6 * 129 + 7 + 7 + 1 + 1 + 3 + 1 + 4 + 2; }
int ref_ns_trace(int fs, int mode, int nframes, const int16_t* pcm_in, float* out_f32, float* trace) {
  if (!(fs == 8000 || fs == 16000)) return -1;
  NsHandle* h = NULL;
  if (WebRtcNs_Create(&h) != 0) return -1;
  if (WebRtcNs_Init(h, (uint32_t)fs) != 0 || WebRtcNs_set_policy(h, mode) != 0) {
    WebRtcNs_Free(h);
    return -1;
  }
  const NoiseSuppressionC* st = (const NoiseSuppressionC*)h;
  const int n = FrameLen(fs), W = ref_ns_trace_words(), ml = st->magnLen;
  std::vector<float> in(n), out(n);
  for (int f = 0; f < nframes; ++f) {
    for (int i = 0; i < n; ++i) in[i] = (float)pcm_in[(size_t)f * n + i];
    const float* inb[1] = {in.data()};
    float* outb[1] = {out.data()};
    WebRtcNs_Analyze(h, in.data());
    WebRtcNs_Process(h, inb, 1, outb);
    if (out_f32) memcpy(out_f32 + (size_t)f * n, out.data(), sizeof(float) * n);
    float* t = trace + (size_t)f * W;
    memset(t, 0, sizeof(float) * W);
    for (int s = 0; s < 3; ++s)
      for (int i = 0; i < ml; ++i) {
        t[s * 129 + i] = st->lquantile[s * ml + i];
        t[387 + s * 129 + i] = st->density[s * ml + i];
      }
    const float* per_bin[6] = {st->quantile, st->smooth, st->noisePrev, st->magnPrevAnalyze, st->logLrtTimeAvg, st->magnAvgPause};
    for (int a = 0; a < 6; ++a) memcpy(t + 774 + a * 129, per_bin[a], sizeof(float) * ml);
    float* u = t + 774 + 6 * 129;
    memcpy(u, st->featureData, sizeof(float) * 7);
    memcpy(u + 7, st->priorModelPars, sizeof(float) * 7);
    u[14] = st->priorSpeechProb;
    u[15] = (float)st->blockInd;
    for (int s = 0; s < 3; ++s) u[16 + s] = (float)st->counter[s];
    u[19] = (float)st->updates;
    for (int k = 0; k < 4; ++k) u[20 + k] = (float)st->modelUpdatePars[k];
    u[24] = st->signalEnergy;
    u[25] = st->sumMagn;
  }
  WebRtcNs_Free(h);
  return 0;
}

// Float NS frame by frame with WebRtcNs_prior_speech_probability read BETWEEN Analyze and Process
// (noise_suppression.c:57-66; 8/16 kHz).  prior_mid: one value per frame.
int ref_ns_prior_between(int fs, int mode, int nframes, const int16_t* pcm_in, float* out_f32, float* prior_mid) {
  if (!(fs == 8000 || fs == 16000)) return -1;
  NsHandle* h = NULL;
  if (WebRtcNs_Create(&h) != 0) return -1;
  if (WebRtcNs_Init(h, (uint32_t)fs) != 0 || WebRtcNs_set_policy(h, mode) != 0) {
    WebRtcNs_Free(h);
    return -1;
  }
  const int n = FrameLen(fs);
  std::vector<float> in(n), out(n);
  for (int f = 0; f < nframes; ++f) {
    for (int i = 0; i < n; ++i) in[i] = (float)pcm_in[(size_t)f * n + i];
    const float* inb[1] = {in.data()};
    float* outb[1] = {out.data()};
    WebRtcNs_Analyze(h, in.data());
    prior_mid[f] = WebRtcNs_prior_speech_probability(h);
    WebRtcNs_Process(h, inb, 1, outb);
    memcpy(out_f32 + (size_t)f * n, out.data(), sizeof(float) * n);
  }
  WebRtcNs_Free(h);
  return 0;
}

// The reference's real FFT itself (utility/fft4g.c:324) with fresh work arrays, as ns_core.c:886-944 calls it.
void ref_rdft(int n, int isgn, float* a) {
  std::vector<int> ip(2 + 64, 0);        // IP_LENGTH of ns_core.h covers n <= 256
  std::vector<float> w(256, 0.f);
  WebRtc_rdft(n, isgn, a, ip.data(), w.data());
}

// ---- primitive hooks for known-answer tests ---------------------------------
// state: 4 x 6 int32 = analysis_state1, analysis_state2, synthesis_state1, synthesis_state2
void ref_qmf_analysis(const int16_t* in, int in_len, int16_t* low, int16_t* high, int32_t* state) {
  WebRtcSpl_AnalysisQMF(in, in_len, low, high, state, state + 6);
}
void ref_qmf_synthesis(const int16_t* low, const int16_t* high, int band_len, int16_t* out,
                       int32_t* state) {
  WebRtcSpl_SynthesisQMF(low, high, band_len, out, state + 12, state + 18);
}

void* ref_resampler_create(int src_frames, int dst_frames) {
  return new webrtc::PushSincResampler(src_frames, dst_frames);
}
void ref_resampler_free(void* r) { delete (webrtc::PushSincResampler*)r; }
int ref_resampler_run(void* r, const int16_t* src, int src_frames, int16_t* dst, int dst_cap) {
  return ((webrtc::PushSincResampler*)r)->Resample(src, src_frames, dst, dst_cap);
}

// Band split / merge exactly as AudioBuffer does it (persistent object so the
// filter states carry across frames).
void* ref_split_create(int fs) {
  const int n = FrameLen(fs);
  return new AudioBuffer(n, 1, n, 1, n);
}
void ref_split_free(void* p) { delete (AudioBuffer*)p; }
// bands: [3][160] int16 (unused bands untouched). Returns num_bands.
int ref_split_analysis(void* p, const int16_t* in, int n, int16_t* bands) {
  AudioBuffer* ab = (AudioBuffer*)p;
  memcpy(ab->data(0), in, sizeof(int16_t) * n);
  ab->SplitIntoFrequencyBands();
  for (int b = 0; b < ab->num_bands(); ++b)
    memcpy(bands + 160 * b, ab->split_bands_const(0)[b], sizeof(int16_t) * 160);
  return ab->num_bands();
}
void ref_split_synthesis(void* p, const int16_t* bands, int n, int16_t* out) {
  AudioBuffer* ab = (AudioBuffer*)p;
  for (int b = 0; b < ab->num_bands(); ++b)
    memcpy(ab->split_bands(0)[b], bands + 160 * b, sizeof(int16_t) * 160);
  ab->MergeFrequencyBands();
  memcpy(out, ab->data_const(0), sizeof(int16_t) * n);
}

// The author's wrapper class (libapm/src/apm_ns.cpp), one AudioBuffer frame per call.
void* ref_apm_ns_create(unsigned fs, int mode, int frame, int channels) {
  APM_NS* a = new APM_NS();
  if (!a->initNsModule(fs, mode, frame, channels)) {
    delete a;
    return NULL;
  }
  return a;
}
void ref_apm_ns_free(void* a) { delete (APM_NS*)a; }
void ref_apm_ns_process_i16(void* a, int16_t* data, int frame, int channels) {
  ((APM_NS*)a)->processCaptureStream(data, frame, channels);
}
void ref_apm_ns_process_f32(void* a, float* data, int frame, int channels) {
  ((APM_NS*)a)->processCaptureStream(data, frame, channels);
}

// SPL scalar helpers (signal_processing_unittest.cc KAT values).
int32_t ref_spl_sqrt_floor(int32_t v) { return WebRtcSpl_SqrtFloor(v); }
int32_t ref_spl_energy(int16_t* v, int len, int* scale) { return WebRtcSpl_Energy(v, len, scale); }
int16_t ref_spl_norm_w32(int32_t a) { return WebRtcSpl_NormW32(a); }
int16_t ref_spl_norm_u32(uint32_t a) { return WebRtcSpl_NormU32(a); }
int16_t ref_spl_norm_w16(int16_t a) { return WebRtcSpl_NormW16(a); }
int16_t ref_spl_get_size_in_bits(uint32_t a) { return WebRtcSpl_GetSizeInBits(a); }
int32_t ref_spl_div_w32w16(int32_t n, int16_t d) { return WebRtcSpl_DivW32W16(n, d); }
uint32_t ref_spl_div_u32u16(uint32_t n, uint16_t d) { return WebRtcSpl_DivU32U16(n, d); }
int16_t ref_spl_max_abs_w16(const int16_t* v, int len) { return WebRtcSpl_MaxAbsValueW16(v, len); }

// order-`order` real FFT pair on int16 (real_fft.c:47,74). fwd: in[1<<order]
// -> out[(1<<order)+2]; inv: in[(1<<order)+2] -> out[1<<order], returns scale.
int ref_spl_real_fft(int order, int inverse, const int16_t* in, int16_t* out) {
  WebRtcSpl_Init();
  struct RealFFT* f = WebRtcSpl_CreateRealFFT(order);
  if (!f) return -1;
  int r = inverse ? WebRtcSpl_RealInverseFFT(f, in, out) : WebRtcSpl_RealForwardFFT(f, in, out);
  WebRtcSpl_FreeRealFFT(f);
  return r;
}

}  // extern "C"
