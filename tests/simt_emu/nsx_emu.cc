// TEST TOOL ONLY: runs the fixed-point NSx CUDA kernel source on the SIMT emulator.
#include "cuda_emu.h"

#include <stdlib.h>

#include <vector>

#include "../../audiosignalprocess_b200/csrc/nsx_host_init.h"
#include "../../audiosignalprocess_b200/csrc/nsx_kernel.cuh"

namespace nsb200 {
uint4 nsx_smem4[(kNsxCtaTableWords + kNsxWarpsPerCta * kNsxWarpWords) / 4 + 4];
}

namespace {
using namespace nsb200;
template <int ANA, int NB>
void Tramp(void* a) { nsx_process_kernel<ANA, NB>(*(const NsxLaunch*)a); }
typedef void (*Fn)(void*);
Fn Pick(int ana, int nb) {
  if (ana == 256 && nb == 1) return (Fn)Tramp<256, 1>;
  if (ana == 256 && nb == 2) return (Fn)Tramp<256, 2>;
  if (ana == 256 && nb == 3) return (Fn)Tramp<256, 3>;
  if (ana == 128 && nb == 1) return (Fn)Tramp<128, 1>;
  return NULL;
}
}  // namespace

extern "C" int emu_nsx_run(int fs, int mode, int nb, int nstreams, int nframes, int fpl,
                           const int16_t* in, int16_t* out) {
  const int ana = fs == 8000 ? 128 : 256;
  const int fl = fs == 8000 ? 80 : 160;
  Fn fn = Pick(ana, nb);
  if (!fn) return -1;
  NsxTables tables;
  nsx_fill_tables(&tables);
  std::vector<uint32_t> state((size_t)nstreams * kNsxStateWords);
  std::vector<int> slots(nstreams);
  for (int s = 0; s < nstreams; ++s) {
    slots[s] = nstreams - 1 - s;
    nsx_init_state(&state[(size_t)slots[s] * kNsxStateWords], (uint32_t)fs);
    nsx_set_mode(&state[(size_t)slots[s] * kNsxStateWords], mode);
  }
  for (int f0 = 0; f0 < nframes; f0 += fpl) {
    NsxLaunch p;
    p.state = state.data();
    p.slots = slots.data();
    p.tables = &tables;
    p.in = in + (size_t)f0 * nb * fl;
    p.out = out + (size_t)f0 * nb * fl;
    p.in_stream_stride = p.out_stream_stride = (long long)nframes * nb * fl;
    p.in_frame_stride = p.out_frame_stride = (long long)nb * fl;
    p.in_band_stride = p.out_band_stride = fl;
    p.n_streams = nstreams;
    p.frames = nframes - f0 < fpl ? nframes - f0 : fpl;
    simt_emu::launch(fn, &p, (nstreams + kNsxWarpsPerCta - 1) / kNsxWarpsPerCta, kNsxWarpsPerCta * 32);
  }
  return 0;
}

// Dumps the product's NSx tables (nsx_host_init.h) for tests/test_tables.py.
extern "C" void emu_nsx_tables(int16_t* win256, int16_t* win128, uint32_t* tw, int16_t* log_frac,
                               int16_t* counter_div, int16_t* log_tab, int16_t* log_idx, int16_t* factor1,
                               int16_t* factor2, int16_t* indicator, int16_t* misc5) {
  nsb200::NsxTables t;
  nsb200::nsx_fill_tables(&t);
  memcpy(win256, t.win256, sizeof(t.win256));
  memcpy(win128, t.win128, sizeof(t.win128));
  memcpy(tw, t.tw, sizeof(t.tw));
  memcpy(log_frac, t.log_frac, sizeof(int16_t) * 256);
  memcpy(counter_div, t.counter_div, sizeof(int16_t) * 201);
  memcpy(log_tab, t.log_tab, sizeof(int16_t) * 9);
  memcpy(log_idx, t.log_idx, sizeof(int16_t) * 129);
  memcpy(factor1, t.factor1, sizeof(int16_t) * 257);
  for (int k = 0; k < 3; ++k) memcpy(factor2 + 257 * k, t.factor2[k], sizeof(int16_t) * 257);
  memcpy(indicator, t.indicator, sizeof(int16_t) * 17);
  misc5[0] = t.sum_log_idx5; misc5[1] = t.sum_sq_log_idx5; misc5[2] = t.det5;
  misc5[3] = t.sum_log_idx65; misc5[4] = t.sum_sq_log_idx65;
}
