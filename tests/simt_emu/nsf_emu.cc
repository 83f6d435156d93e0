// TEST TOOL ONLY: runs the float-NS CUDA kernel source on the SIMT emulator.
#include "cuda_emu.h"

#include <stdlib.h>

#include <vector>

#include "../../audiosignalprocess_b200/csrc/nsf_host_init.h"
#include "../../audiosignalprocess_b200/csrc/nsf_kernel.cuh"

namespace nsb200 {
float4 nsf_smem4[(kNsfCtaTableWords + kNsfWarpsPerCta * kNsfWarpWordsMax) / 4 + 4];
}

namespace {
using namespace nsb200;

template <int ANA, int NB, bool I16, bool SPLIT = false>
void Tramp(void* a) { nsf_process_kernel<ANA, NB, I16, SPLIT>(*(const NsfLaunch*)a); }

typedef void (*Fn)(void*);
Fn Pick(int ana, int nb, bool i16) {
#define C(A, N) if (ana == A && nb == N) return i16 ? (Fn)Tramp<A, N, true> : (Fn)Tramp<A, N, false>;
  C(256, 1) C(256, 2) C(256, 3) C(128, 1)
#undef C
  return NULL;
}
Fn PickSplit(int ana, int nb) {
  if (ana == 256 && nb == 1) return (Fn)Tramp<256, 1, false, true>;
  if (ana == 256 && nb == 2) return (Fn)Tramp<256, 2, false, true>;
  if (ana == 128 && nb == 1) return (Fn)Tramp<128, 1, false, true>;
  return NULL;
}
}  // namespace

extern "C" {

// Streams laid out [stream][frame][band][frame_len]; in/out int16 (i16=1) or
// float.  fpl = frames per emulated launch (state round-trips the slab between
// launches exactly as on the device).  prior_prob: [stream][frame] or NULL
// (only filled when fpl == 1).
// state_out (optional): [stream][kNsfStateWords] words, the slab of every stream after the last frame.
int emu_nsf_run_state(int fs, int mode, int nb, int i16, int nstreams, int nframes, int fpl,
                      const void* in, void* out, float* prior_prob, uint32_t* state_out) {
  const int ana = fs == 8000 ? 128 : 256;
  const int fl = fs == 8000 ? 80 : 160;
  Fn fn = Pick(ana, nb, i16 != 0);
  if (!fn) return -1;
  NsfTables tables;
  nsf_fill_tables(&tables);
  std::vector<uint32_t> state((size_t)nstreams * kNsfStateWords);
  std::vector<int> hist((size_t)nstreams * kNsfHistWords, 0);
  std::vector<int> slots(nstreams);
  for (int s = 0; s < nstreams; ++s) {
    slots[s] = nstreams - 1 - s;  // exercise the indirection
    nsf_init_state(&state[(size_t)slots[s] * kNsfStateWords], (uint32_t)fs);
    nsf_set_mode(&state[(size_t)slots[s] * kNsfStateWords], mode);
  }
  const size_t esz = i16 ? 2 : 4;
  for (int f0 = 0; f0 < nframes; f0 += fpl) {
    NsfLaunch p;
    p.state = (float*)state.data();
    p.hist = hist.data();
    p.slots = slots.data();
    p.tables = &tables;
    p.in = (const char*)in + (size_t)f0 * nb * fl * esz;
    p.out = (char*)out + (size_t)f0 * nb * fl * esz;
    p.in_stream_stride = p.out_stream_stride = (long long)nframes * nb * fl;
    p.in_frame_stride = p.out_frame_stride = (long long)nb * fl;
    p.in_band_stride = p.out_band_stride = fl;
    p.n_streams = nstreams;
    p.frames = nframes - f0 < fpl ? nframes - f0 : fpl;
    simt_emu::launch(fn, &p, (nstreams + kNsfWarpsPerCta - 1) / kNsfWarpsPerCta, kNsfWarpsPerCta * 32);
    if (prior_prob && fpl == 1)
      for (int s = 0; s < nstreams; ++s)
        prior_prob[(size_t)s * nframes + f0] =
            ((float*)&state[(size_t)slots[s] * kNsfStateWords])[kH_priorSpeechProb];
  }
  if (state_out)
    for (int s = 0; s < nstreams; ++s)
      memcpy(state_out + (size_t)s * kNsfStateWords, &state[(size_t)slots[s] * kNsfStateWords], sizeof(uint32_t) * kNsfStateWords);
  return 0;
}
int emu_nsf_run(int fs, int mode, int nb, int i16, int nstreams, int nframes, int fpl,
                const void* in, void* out, float* prior_prob) {
  return emu_nsf_run_state(fs, mode, nb, i16, nstreams, nframes, fpl, in, out, prior_prob, NULL);
}
int emu_nsf_state_words(void) { return kNsfStateWords; }

// Split mode (float samples): Analyze sees ana[stream][frame][fl], Process sees in[stream][frame][band][fl].
// The first `fused_frames` frames run through the fused kernel on `in` (ana ignored) so that the
// fused -> split hand-over of a running stream is exercised.
// phased != 0: every launch of the split kernel is issued as two, the Analyze half then the Process half (the
// single-stream WebRtcNs_Analyze / WebRtcNs_Process pair); prior_mid (optional, [stream][frame], fpl == 1):
// the prior speech probability in the state between the two.
int emu_nsf_run_split_phased(int fs, int mode, int nb, int nstreams, int nframes, int fpl, int fused_frames,
                             const float* ana_in, const float* in, float* out, int phased, float* prior_mid) {
  const int ana = fs == 8000 ? 128 : 256;
  const int fl = fs == 8000 ? 80 : 160;
  Fn fs_fn = PickSplit(ana, nb), ff_fn = Pick(ana, nb, false);
  if (!fs_fn || !ff_fn) return -1;
  NsfTables tables;
  nsf_fill_tables(&tables);
  std::vector<uint32_t> state((size_t)nstreams * kNsfStateWords);
  std::vector<int> hist((size_t)nstreams * kNsfHistWords, 0);
  std::vector<int> slots(nstreams);
  for (int s = 0; s < nstreams; ++s) {
    slots[s] = s;
    nsf_init_state(&state[(size_t)s * kNsfStateWords], (uint32_t)fs);
    nsf_set_mode(&state[(size_t)s * kNsfStateWords], mode);
  }
  for (int f0 = 0; f0 < nframes;) {
    const bool fused = f0 < fused_frames;
    int nf = nframes - f0 < fpl ? nframes - f0 : fpl;
    if (fused && f0 + nf > fused_frames) nf = fused_frames - f0;
    NsfLaunch p;
    p.state = (float*)state.data();
    p.hist = hist.data();
    p.slots = slots.data();
    p.tables = &tables;
    p.in = in + (size_t)f0 * nb * fl;
    p.out = out + (size_t)f0 * nb * fl;
    p.in_stream_stride = p.out_stream_stride = (long long)nframes * nb * fl;
    p.in_frame_stride = p.out_frame_stride = (long long)nb * fl;
    p.in_band_stride = p.out_band_stride = fl;
    p.ana_in = ana_in + (size_t)f0 * fl;
    p.ana_stream_stride = (long long)nframes * fl;
    p.ana_frame_stride = fl;
    p.n_streams = nstreams;
    p.frames = nf;
    const int grid = (nstreams + kNsfWarpsPerCta - 1) / kNsfWarpsPerCta;
    if (fused || !phased) {
      simt_emu::launch(fused ? ff_fn : fs_fn, &p, grid, kNsfWarpsPerCta * 32);
    } else {
      p.phase = 1;
      simt_emu::launch(fs_fn, &p, grid, kNsfWarpsPerCta * 32);
      if (prior_mid && fpl == 1)
        for (int s = 0; s < nstreams; ++s)
          prior_mid[(size_t)s * nframes + f0] = ((float*)&state[(size_t)s * kNsfStateWords])[kH_priorSpeechProb];
      p.phase = 2;
      simt_emu::launch(fs_fn, &p, grid, kNsfWarpsPerCta * 32);
    }
    f0 += nf;
  }
  return 0;
}
int emu_nsf_run_split(int fs, int mode, int nb, int nstreams, int nframes, int fpl, int fused_frames,
                      const float* ana_in, const float* in, float* out) {
  return emu_nsf_run_split_phased(fs, mode, nb, nstreams, nframes, fpl, fused_frames, ana_in, in, out, 0, NULL);
}

void emu_nsf_tables(float* win256, float* win128) {
  NsfTables t;
  nsf_fill_tables(&t);
  memcpy(win256, t.win256, sizeof(t.win256));
  memcpy(win128, t.win128, sizeof(t.win128));
}

}  // extern "C"

#include "../../audiosignalprocess_b200/csrc/pcm_synth.h"
extern "C" void emu_pcm_synth(uint32_t seed, uint32_t stream, uint32_t fs, uint32_t n0, uint32_t n,
                              int16_t* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = pcm_synth_sample(seed, stream, fs, n0 + i);
}
