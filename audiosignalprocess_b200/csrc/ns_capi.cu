// C-ABI host layer of libwebrtc_ns_b200.so (include/webrtc_ns_b200.h).
//
// Mirrors the reference's thin C wrappers (ns/noise_suppression.c:20-66,
// ns/noise_suppression_x.c:19-54): a handle is a small host struct that names a
// slot in a per-GPU state slab; Create/Init/set_policy edit that slot, the
// batch calls hand a list of slots to one kernel launch.  No CPU compute path
// exists here: without a usable GPU every call fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/webrtc_ns_b200.h"
#include "nsf_host_init.h"
#include "nsf_kernel.cuh"
#include "nsx_host_init.h"
#include "nsx_kernel.cuh"
#include "band_host_init.h"
#include "band_kernels.cuh"
#include "pcm_synth.h"

namespace nsb200 {
namespace {

// Threading.  The reference keeps no global state at all (ns/noise_suppression.c:20-66: every call works on
// its own handle), so independent handles may be driven from independent threads.  Here handles share per-GPU
// slabs, streams and staging buffers, so the rule is per device: calls that only USE devices (every batch /
// process / getter call) hold g_rw shared plus the mutex of each device they touch -- two host threads driving
// two GPUs run concurrently, GPU waits included -- and calls that change what handles exist or where they live
// (Create, Free, Init, migration, state import) hold g_rw exclusively.  The last error, the device new handles
// are created on and the validated-handle-list memo are per thread.
std::shared_mutex g_rw;
thread_local std::string t_err;
std::atomic<uint64_t> g_launches{0};
thread_local int t_create_device = -1;

int Fail(const std::string& m) {
  t_err = m;
  return -1;
}
#define CU_OK(call)                                                              \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess)                                                       \
      return Fail(std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

constexpr uint32_t kMagicF = 0x4e53464cu;  // float handle
constexpr uint32_t kMagicX = 0x4e535846u;  // fixed handle

struct Handle {
  uint32_t magic;
  int dev;
  int slot;
  uint32_t fs;
  int mode;
  int init_flag;
  bool analyze_seen;
  float analyze_frame[160];
  uint64_t seen_stamp;   // batch validation: the call that last listed this handle (duplicates race on one slab)
  bool split_mode;   // float NS: the stream has been fed distinct Analyze / Process signals (sticky)
  double down_vsi;   // 48 kHz: running position of the 640 -> 480 resampler (band_host_init.h)
};

// Growable pool of fixed-size slabs in device memory.
struct SlabPool {
  void* base = nullptr;
  size_t slab_bytes = 0;
  int capacity = 0;
  std::vector<int> free_slots;
  int next = 0;
};

struct DeviceCtx {
  std::mutex mu;                      // held by every call that uses this device (see "Threading")
  int dev = -1;
  bool ready = false;
  cudaEvent_t user_done = nullptr;    // last launch a *Device entry point enqueued on a caller's stream
  cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
  NsfTables* d_nsf_tables = nullptr;
  NsxTables* d_nsx_tables = nullptr;
  float* d_sinc_up = nullptr;     // 33 x 32 taps, 480 -> 640
  float* d_sinc_down = nullptr;   // 33 x 32 taps, 640 -> 480
  std::vector<std::vector<uint8_t>> down_regular;   // scratch of RunBandBlock (48 kHz merge)
  SlabPool f_state, f_hist, x_state, b_state;  // float state, float histograms, fixed state, band-split state
  void* d_template = nullptr;   // scratch for Init templates
  size_t template_bytes = 0;
  int* d_slots = nullptr;       // slot list of the current batch
  int* h_slots = nullptr;       // pinned
  int slots_cap = 0;
  std::vector<int> cached_slots;
  int slot_run = 0;             // cached_slots[i] == cached_slots[0] + i for i < slot_run
  // device staging for the host-pointer batch API
  int16_t* d_in = nullptr;
  int16_t* d_out = nullptr;
  size_t stage_elems = 0;
  int16_t* d_bands = nullptr;   // [stream][frame][band][160] scratch for 32/48 kHz
  size_t bands_elems = 0;
  float* d_single = nullptr;           // single-frame float calls (WebRtcNs_Analyze / WebRtcNs_Process): in | out | ana
  int16_t* d_band_scratch = nullptr;   // 48 kHz: 64 kHz and 32 kHz intermediates
  size_t band_scratch_elems = 0;
  int32_t* d_down_sched = nullptr;     // 48 kHz: resampler schedule of the current launch
  size_t down_sched_words = 0;
  std::vector<cudaEvent_t> events;     // host-pointer batch pipeline (reused across calls)
  // staging-buffer hand-over that survives the call (asynchronous batches form one pipeline):
  // buffer b was last read by the kernel before pipe_kdone[b] and drained before pipe_drained[b]
  cudaEvent_t pipe_kdone[8] = {}, pipe_drained[8] = {};
  bool pipe_used[8] = {};
  int pipe_bufs = 0;
  size_t pipe_per = 0;                 // staging layout of the pipeline: samples per stream per buffer,
  int pipe_count = 0;                  //   streams per buffer
  unsigned long long pipe_seq = 0;     // chunks issued so far (buffer = seq % 3)
  // 32/48 kHz: one CUDA stream per pipeline stage, events between the stages of a chunk
  cudaStream_t stage_stream[9] = {}, stage_stream_hi[9] = {};   // per pipeline position; _hi: the QMF stages
  cudaEvent_t fork_event = nullptr;
  std::vector<cudaEvent_t> band_events;
};

// (DeviceCtx holds a mutex: a fixed array, sized once)
struct DeviceList {
  std::unique_ptr<DeviceCtx[]> p;
  int n = 0;
  size_t size() const { return (size_t)n; }
  bool empty() const { return n == 0; }
  DeviceCtx& operator[](size_t i) { return p[i]; }
  DeviceCtx* begin() { return p.get(); }
  DeviceCtx* end() { return p.get() + n; }
};
DeviceList g_devs;
std::once_flag g_devs_once;

// Asynchronous host-pointer batches (WebRtcNs[x]_ProcessBatchAsync) still in flight: ticket ->
// one completion event per device.  Every other call that uses a device first waits for the tickets that
// involve it (UseLock::Device), so the rest of the library keeps its "nothing of mine is running on this
// device" assumption.
struct PendingBatch {
  uint64_t ticket;
  std::vector<std::pair<int, cudaEvent_t>> done;
};
std::mutex g_pend_mu;
std::vector<PendingBatch> g_pending;
uint64_t g_next_ticket = 1;

void FinishPending(PendingBatch& p) {
  for (auto& de : p.done) {
    cudaSetDevice(de.first);
    cudaEventSynchronize(de.second);
    cudaEventDestroy(de.second);
  }
  p.done.clear();
}
// dev < 0: every ticket
void DrainPending(int dev) {
  std::vector<PendingBatch> mine;
  {
    std::lock_guard<std::mutex> g(g_pend_mu);
    for (size_t i = 0; i < g_pending.size();) {
      bool hit = dev < 0;
      for (auto& de : g_pending[i].done) hit = hit || de.first == dev;
      if (hit) {
        mine.push_back(std::move(g_pending[i]));
        g_pending.erase(g_pending.begin() + (long)i);
      } else {
        ++i;
      }
    }
  }
  for (PendingBatch& p : mine) FinishPending(p);
}
// true between CheckBatch and the end of the API call: the call's handle list IS t_memo.hs, so the
// helpers below need not compare 8 bytes per stream again (a one-frame tick over 32768 streams is
// ~125 us of GPU time; the host side of the call has to stay well below that)
thread_local bool t_call_is_memo = false;
// The library switches the calling thread's current device as it works through a call's devices; the caller
// gets its own back (it decides where handles are created when WebRtcNsB200_SetCreateDevice was not used).
struct KeepCallerDevice {
  int saved = -1;
  KeepCallerDevice() {
    if (cudaGetDevice(&saved) != cudaSuccess) {
      saved = -1;
      cudaGetLastError();
    }
  }
  ~KeepCallerDevice() {
    if (saved >= 0) cudaSetDevice(saved);
  }
};
// Changes which handles exist or where they live: alone in the library, nothing in flight.
struct ExclusiveLock {
  KeepCallerDevice keep;
  std::unique_lock<std::shared_mutex> guard;
  ExclusiveLock() : guard(g_rw) {
    t_call_is_memo = false;
    DrainPending(-1);
  }
};
// Uses devices: shared with other users, one mutex per device touched (taken in ascending device order
// when a call spans several).
struct UseLock {
  KeepCallerDevice keep;
  std::shared_lock<std::shared_mutex> guard;
  std::vector<std::unique_lock<std::mutex>> held;
  UseLock() : guard(g_rw) { t_call_is_memo = false; }
  void Device(int dev, bool drain = true);
};

int EnsureDevices() {
  std::call_once(g_devs_once, [] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      t_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
      return;
    }
    g_devs.p.reset(new DeviceCtx[n]);
    g_devs.n = n;
    for (int i = 0; i < n; ++i) g_devs[i].dev = i;
  });
  if (g_devs.empty()) {
    if (t_err.empty() || t_err.compare(0, 14, "no CUDA device") != 0) {
      int n = 0;
      cudaError_t e = cudaGetDeviceCount(&n);
      return Fail(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0"));
    }
    return -1;
  }
  return 0;
}

void UseLock::Device(int dev, bool drain) {
  if (EnsureDevices() != 0 || dev < 0 || dev >= (int)g_devs.size()) return;   // DeviceReady reports it
  held.emplace_back(g_devs[dev].mu);
  if (drain) DrainPending(dev);
}

int PoolGrow(DeviceCtx& d, SlabPool& p, int want) {
  if (want <= p.capacity) return 0;
  int cap = p.capacity ? p.capacity : 1024;
  while (cap < want) cap *= 2;
  void* nb = nullptr;
  CU_OK(cudaMalloc(&nb, (size_t)cap * p.slab_bytes));
  if (p.base) {
    CU_OK(cudaDeviceSynchronize());
    CU_OK(cudaMemcpy(nb, p.base, (size_t)p.capacity * p.slab_bytes, cudaMemcpyDeviceToDevice));
    CU_OK(cudaFree(p.base));
  }
  p.base = nb;
  p.capacity = cap;
  return 0;
}

int PoolAlloc(DeviceCtx& d, SlabPool& p, int* slot) {
  if (!p.free_slots.empty()) {
    *slot = p.free_slots.back();
    p.free_slots.pop_back();
    return 0;
  }
  if (p.next >= p.capacity && PoolGrow(d, p, p.next + 1) != 0) return -1;
  *slot = p.next++;
  return 0;
}

int DeviceReady(int dev, DeviceCtx** out) {
  if (EnsureDevices() != 0) return -1;
  if (dev < 0 || dev >= (int)g_devs.size()) return Fail("bad device index");
  DeviceCtx& d = g_devs[dev];
  CU_OK(cudaSetDevice(dev));
  if (!d.ready) {
    CU_OK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CU_OK(cudaStreamCreateWithFlags(&d.copy_in, cudaStreamNonBlocking));
    CU_OK(cudaStreamCreateWithFlags(&d.copy_out, cudaStreamNonBlocking));
    {
      NsfTables* t = new NsfTables;
      nsf_fill_tables(t);
      CU_OK(cudaMalloc(&d.d_nsf_tables, sizeof(NsfTables)));
      CU_OK(cudaMemcpy(d.d_nsf_tables, t, sizeof(NsfTables), cudaMemcpyHostToDevice));
      delete t;
    }
    {
      NsxTables* t = new NsxTables;
      nsx_fill_tables(t);
      CU_OK(cudaMalloc(&d.d_nsx_tables, sizeof(NsxTables)));
      CU_OK(cudaMemcpy(d.d_nsx_tables, t, sizeof(NsxTables), cudaMemcpyHostToDevice));
      delete t;
    }
    {
      std::vector<float> k(33 * 32);
      band_make_sinc_kernel(480.0 / 640.0, k.data());
      CU_OK(cudaMalloc(&d.d_sinc_up, sizeof(float) * k.size()));
      CU_OK(cudaMemcpy(d.d_sinc_up, k.data(), sizeof(float) * k.size(), cudaMemcpyHostToDevice));
      {
        // the 480 -> 640 resampler only ever uses table rows 16, 8, 0, 24 (band_kernels.cuh)
        float rows[4][32];
        const int which[4] = {0, 8, 16, 24};
        for (int r = 0; r < 4; ++r) memcpy(rows[r], k.data() + which[r] * 32, sizeof(rows[r]));
        CU_OK(cudaMemcpyToSymbol(c_up_rows, rows, sizeof(rows)));
        float2 pairs[2][4][16];
        for (int r = 0; r < 4; ++r)
          for (int kk = 0; kk < 16; ++kk) {
            pairs[0][r][kk] = make_float2(rows[r][2 * kk], rows[r][2 * kk + 1]);
            pairs[1][r][kk] = kk < 15 ? make_float2(rows[r][2 * kk + 1], rows[r][2 * kk + 2]) : make_float2(0.f, 0.f);
          }
        CU_OK(cudaMemcpyToSymbol(c_up_pairs, pairs, sizeof(pairs)));
      }
      band_make_sinc_kernel(640.0 / 480.0, k.data());
      CU_OK(cudaMalloc(&d.d_sinc_down, sizeof(float) * k.size()));
      CU_OK(cudaMemcpy(d.d_sinc_down, k.data(), sizeof(float) * k.size(), cudaMemcpyHostToDevice));
      {
        // the rows a regular 640 -> 480 schedule uses, as (row, row + 1) pairs per tap (band_kernels.cuh)
        float2 pairs[3][32];
        for (int ph = 0; ph < 3; ++ph)
          for (int i = 0; i < 32; ++i)
            pairs[ph][i] = make_float2(k[(size_t)kDownRegularRows[ph] * 32 + i], k[(size_t)(kDownRegularRows[ph] + 1) * 32 + i]);
        CU_OK(cudaMemcpyToSymbol(c_down_pairs, pairs, sizeof(pairs)));
      }
    }
    d.f_state.slab_bytes = sizeof(uint32_t) * kNsfStateWords;
    d.f_hist.slab_bytes = sizeof(uint32_t) * kNsfHistWords;
    d.x_state.slab_bytes = sizeof(uint32_t) * kNsxStateWords;
    d.b_state.slab_bytes = sizeof(uint32_t) * kBandStateWords;
    {
      int w = kNsfStateWords > kNsxStateWords ? kNsfStateWords : kNsxStateWords;
      if (kBandStateWords > w) w = kBandStateWords;
      d.template_bytes = sizeof(uint32_t) * w;
    }
    CU_OK(cudaMalloc(&d.d_template, d.template_bytes));
    d.ready = true;
  }
  *out = &d;
  return 0;
}

// A caller that ticks the same handle list every 10 ms (F = 1) should not pay for re-validating
// tens of thousands of handles per call: the last validated list is remembered, and anything that
// could invalidate it (Create / Free / Init / migration / import / a stream turning two-signal)
// bumps g_epoch.  Measured at 32 768 streams: ~340 us of host work per call against a 220 us kernel.
std::atomic<uint64_t> g_epoch{1};
std::atomic<uint64_t> g_call_stamp{1};
struct BatchMemo {
  uint64_t epoch = 0;
  uint32_t magic = 0;
  std::vector<void*> hv;        // the caller's list
  std::vector<Handle*> hs;      // ... validated
  std::vector<int> slots;       // RunDevice's slot list for exactly `hs` (empty = not built)
  bool any_split = false;       // some handle of `hs` is in two-signal mode
  int single_dev = -1;          // the device all of `hs` live on, -1 if they span several
};
thread_local BatchMemo t_memo;   // per calling thread: the list that thread ticks

Handle* AsHandle(void* h, uint32_t magic) {
  Handle* p = static_cast<Handle*>(h);
  return (p && p->magic == magic) ? p : nullptr;
}

// ---- init: copy a template slab into each listed slot, zero the cold slab ----
__global__ void slab_fill_kernel(uint32_t* base, int words, const uint32_t* tmpl, const int* slots,
                                 int n, uint32_t* cold, int cold_words, uint32_t* aux, int aux_words) {
  const int s = blockIdx.x;
  if (s >= n) return;
  const int slot = slots[s];
  uint32_t* dst = base + (size_t)slot * words;
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = tmpl[i];
  if (cold) {
    uint32_t* c = cold + (size_t)slot * cold_words;
    for (int i = threadIdx.x; i < cold_words; i += blockDim.x) c[i] = 0u;
  }
  if (aux) {
    uint32_t* a = aux + (size_t)slot * aux_words;
    for (int i = threadIdx.x; i < aux_words; i += blockDim.x) a[i] = 0u;
  }
}

int UploadSlots(DeviceCtx& d, const std::vector<int>& slots, cudaStream_t st) {
  const int n = (int)slots.size();
  if (n > d.slots_cap) {
    if (d.d_slots) {
      CU_OK(cudaDeviceSynchronize());
      CU_OK(cudaFree(d.d_slots));
      CU_OK(cudaFreeHost(d.h_slots));
    }
    int cap = 1024;
    while (cap < n) cap *= 2;
    CU_OK(cudaMalloc(&d.d_slots, sizeof(int) * cap));
    CU_OK(cudaMallocHost(&d.h_slots, sizeof(int) * cap));
    d.slots_cap = cap;
    d.cached_slots.clear();
  }
  if (d.cached_slots == slots) return 0;  // same batch as last call: list already resident
  // a kernel of the previous batch (on any stream) may still be reading the old list
  CU_OK(cudaDeviceSynchronize());
  memcpy(d.h_slots, slots.data(), sizeof(int) * n);
  CU_OK(cudaMemcpyAsync(d.d_slots, d.h_slots, sizeof(int) * n, cudaMemcpyHostToDevice, d.stream));
  if (st != d.stream) CU_OK(cudaStreamSynchronize(d.stream));
  d.cached_slots = slots;
  // handles created one after the other sit in consecutive slots: the kernel then computes the
  // slot instead of loading it (one dependent global-memory latency less per warp)
  int run = n > 0 ? 1 : 0;
  while (run < n && slots[run] == slots[0] + run) ++run;
  d.slot_run = run;
  return 0;
}

int Create(void** out, uint32_t magic) {
  ExclusiveLock lk;
  ++g_epoch;   // handle lists validated so far are stale (BatchMemo)
  if (!out) return Fail("NULL handle pointer");
  *out = nullptr;
  int dev = t_create_device;
  if (dev < 0) {
    if (EnsureDevices() != 0) return -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return Fail("cudaGetDevice failed");
  }
  DeviceCtx* d;
  if (DeviceReady(dev, &d) != 0) return -1;
  Handle* h = new Handle();
  h->magic = magic;
  h->dev = dev;
  h->init_flag = 0;
  h->fs = 0;
  h->mode = 0;
  h->analyze_seen = false;
  h->split_mode = false;
  int rc;
  if (magic == kMagicF) {
    rc = PoolAlloc(*d, d->f_state, &h->slot);
    if (rc == 0 && PoolGrow(*d, d->f_hist, d->f_state.capacity) != 0) rc = -1;
  } else {
    rc = PoolAlloc(*d, d->x_state, &h->slot);
  }
  if (rc == 0 && PoolGrow(*d, d->b_state, 2 * (magic == kMagicF ? d->f_state : d->x_state).capacity) != 0) rc = -1;
  if (rc != 0) {
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

int Free(void* hv, uint32_t magic) {
  ExclusiveLock lk;
  ++g_epoch;   // handle lists validated so far are stale (BatchMemo)
  Handle* h = AsHandle(hv, magic);
  if (!h) return 0;  // reference: free(NULL) is fine, returns 0
  DeviceCtx& d = g_devs[h->dev];
  cudaSetDevice(h->dev);
  cudaStreamSynchronize(d.stream);
  (magic == kMagicF ? d.f_state : d.x_state).free_slots.push_back(h->slot);
  d.cached_slots.clear();
  h->magic = 0;
  delete h;
  return 0;
}

// Band-split state shares the slot index of the owning (float or fixed) handle;
// float and fixed handles use disjoint halves of the band pool.
int BandSlot(const Handle* h) { return h->magic == kMagicF ? 2 * h->slot : 2 * h->slot + 1; }

int InitMany(void* const* hv, int n, uint32_t fs, int mode, uint32_t magic) {
  ExclusiveLock lk;
  ++g_epoch;   // handle lists validated so far are stale (BatchMemo)
  if (!hv || n <= 0) return Fail("no handles");
  if (!(fs == 8000 || fs == 16000 || fs == 32000 || fs == 48000)) return Fail("unsupported fs");
  if (mode < 0 || mode > 3) return Fail("mode out of range");
  std::vector<std::vector<int>> per_dev(g_devs.size());
  for (int i = 0; i < n; ++i) {
    Handle* h = AsHandle(hv[i], magic);
    if (!h) return Fail("bad handle");
    per_dev[h->dev].push_back(i);
  }
  for (size_t dv = 0; dv < per_dev.size(); ++dv) {
    if (per_dev[dv].empty()) continue;
    DeviceCtx* d;
    if (DeviceReady((int)dv, &d) != 0) return -1;
    std::vector<int> slots, bslots;
    for (int i : per_dev[dv]) {
      Handle* h = static_cast<Handle*>(hv[i]);
      slots.push_back(h->slot);
      bslots.push_back(BandSlot(h));
    }
    std::vector<uint32_t> tmpl(d->template_bytes / 4, 0u);
    int words;
    if (magic == kMagicF) {
      nsf_init_state(tmpl.data(), fs);
      nsf_set_mode(tmpl.data(), mode);
      words = kNsfStateWords;
    } else {
      nsx_init_state(tmpl.data(), fs);
      nsx_set_mode(tmpl.data(), mode);
      words = kNsxStateWords;
    }
    CU_OK(cudaStreamSynchronize(d->stream));
    CU_OK(cudaMemcpyAsync(d->d_template, tmpl.data(), sizeof(uint32_t) * words, cudaMemcpyHostToDevice, d->stream));
    if (PoolGrow(*d, d->b_state, 2 * (magic == kMagicF ? d->f_state : d->x_state).capacity) != 0) return -1;
    if (UploadSlots(*d, slots, d->stream) != 0) return -1;
    const int m = (int)slots.size();
    if (magic == kMagicF) {
      slab_fill_kernel<<<m, 256, 0, d->stream>>>((uint32_t*)d->f_state.base, kNsfStateWords,
                                                 (const uint32_t*)d->d_template, d->d_slots, m,
                                                 (uint32_t*)d->f_hist.base, kNsfHistWords, nullptr, 0);
    } else {
      slab_fill_kernel<<<m, 256, 0, d->stream>>>((uint32_t*)d->x_state.base, kNsxStateWords,
                                                 (const uint32_t*)d->d_template, d->d_slots, m,
                                                 nullptr, 0, nullptr, 0);
    }
    ++g_launches;
    // band-split state: zero filter states (TwoBandsStates ctor, splitting_filter.h:34-39),
    // resamplers as their zero-primed first pass leaves them
    if (UploadSlots(*d, bslots, d->stream) != 0) return -1;
    {
      std::vector<uint32_t> bt(kBandStateWords);
      band_init_state(bt.data());
      CU_OK(cudaMemcpyAsync(d->d_template, bt.data(), sizeof(uint32_t) * kBandStateWords, cudaMemcpyHostToDevice, d->stream));
    }
    slab_fill_kernel<<<m, 256, 0, d->stream>>>((uint32_t*)d->b_state.base, kBandStateWords,
                                               (const uint32_t*)d->d_template, d->d_slots, m,
                                               nullptr, 0, nullptr, 0);
    ++g_launches;
    CU_OK(cudaGetLastError());
    CU_OK(cudaStreamSynchronize(d->stream));
    d->cached_slots.clear();
  }
  for (int i = 0; i < n; ++i) {
    Handle* h = static_cast<Handle*>(hv[i]);
    h->fs = fs;
    h->mode = mode;
    h->init_flag = 1;
    h->analyze_seen = false;
    h->down_vsi = band_init_vsi();
  }
  return 0;
}

int SetPolicy(void* hv, int mode, uint32_t magic) {
  Handle* h;
  {
    UseLock lk;
    h = AsHandle(hv, magic);
    if (!h) return Fail("bad handle");
    if (mode < 0 || mode > 3) return Fail("mode out of range");
    lk.Device(h->dev);
    if (!h->init_flag) {
      // the reference only writes fields of the struct here; Init later resets
      // them to mode 0, so this is a no-op before Init.
      return 0;
    }
    DeviceCtx* d;
    if (DeviceReady(h->dev, &d) != 0) return -1;
    CU_OK(cudaStreamSynchronize(d->stream));
    if (magic == kMagicF) {
      uint32_t w[kNsfHdrWords];
      uint32_t* slab = (uint32_t*)d->f_state.base + (size_t)h->slot * kNsfStateWords;
      CU_OK(cudaMemcpy(w, slab, sizeof(w), cudaMemcpyDeviceToHost));
      // nsf_set_mode touches header words only
      std::vector<uint32_t> tmp(kNsfStateWords, 0u);
      memcpy(tmp.data(), w, sizeof(w));
      nsf_set_mode(tmp.data(), mode);
      CU_OK(cudaMemcpy(slab, tmp.data(), sizeof(w), cudaMemcpyHostToDevice));
    } else {
      uint32_t w[kNsxHdrWords];
      uint32_t* slab = (uint32_t*)d->x_state.base + (size_t)h->slot * kNsxStateWords;
      CU_OK(cudaMemcpy(w, slab, sizeof(w), cudaMemcpyDeviceToHost));
      std::vector<uint32_t> tmp(kNsxStateWords, 0u);
      memcpy(tmp.data(), w, sizeof(w));
      nsx_set_mode(tmp.data(), mode);
      CU_OK(cudaMemcpy(slab, tmp.data(), sizeof(w), cudaMemcpyHostToDevice));
    }
    h->mode = mode;
  }
  return 0;
}

// ---- kernel dispatch ---------------------------------------------------------
// Residency: measured on the B200 (tools/bench_variants.sh, profiles/r1_tuning.md) the float
// kernel is fastest at 8 two-warp CTAs per SM (128 registers, no spills that matter) and the
// fixed-point kernel at 14 (72 registers); padding shared memory to force "balanced" full waves
// (7 CTAs/SM = exactly two rounds for 4096 streams) was slower than the fuller occupancy.
template <int ANA, int NB, bool I16, bool SPLIT>
int LaunchNsfT(const NsfLaunch& p, cudaStream_t st) {
  const int grid = (p.n_streams + kNsfWarpsPerCta - 1) / kNsfWarpsPerCta;
  // batches of several waves: each warp prefetches for the stream one resident wave ahead
  static int sms[64] = {0};
  int dev = 0;
  CU_OK(cudaGetDevice(&dev));
  if (sms[dev & 63] == 0) CU_OK(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  NsfLaunch q = p;
  // (measured at one frame per launch over 32 768 streams: off 127.2 us, 0.06 / 0.125 / 0.25 / 0.375 / 0.5 /
  // 1 / 2 resident waves ahead 127.0 / 122.9 / 120.8 / 119.8 / 120.6 / 122.9 / 131.0 us)
  q.prefetch_ahead = sms[dev & 63] * kNsfCtasPerSm * kNsfWarpsPerCta / 2;
  if (const char* e = getenv("NSB200_NSF_AHEAD")) q.prefetch_ahead = atoi(e);   // tuning: 0 = off
  if (p.frames == 1 && !getenv("NSB200_NSF_NOWRAP")) q.prefetch_ahead = -q.prefetch_ahead;   // tick: wrap to the batch head
  const size_t smem = sizeof(float) * (kNsfCtaTableWords + kNsfWarpsPerCta * NsfWarpWords<SPLIT, NB>::value);
  if (smem > 48 * 1024)
    CU_OK(cudaFuncSetAttribute(nsf_process_kernel<ANA, NB, I16, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nsf_process_kernel<ANA, NB, I16, SPLIT><<<grid, kNsfWarpsPerCta * 32, smem, st>>>(q);
  ++g_launches;
  CU_OK(cudaGetLastError());
  return 0;
}
template <int ANA, int NB>
int LaunchNsfV(bool i16, bool split, const NsfLaunch& p, cudaStream_t st) {
  if (split) return i16 ? LaunchNsfT<ANA, NB, true, true>(p, st) : LaunchNsfT<ANA, NB, false, true>(p, st);
  return i16 ? LaunchNsfT<ANA, NB, true, false>(p, st) : LaunchNsfT<ANA, NB, false, false>(p, st);
}
// split: Analyze is fed p.ana_in, Process p.in (nsf_kernel.cuh, SPLIT)
int LaunchNsf(int ana, int nb, bool i16, bool split, const NsfLaunch& p, cudaStream_t st) {
  if (ana == 128 && nb == 1) return LaunchNsfV<128, 1>(i16, split, p, st);
  if (ana == 256 && nb == 1) return LaunchNsfV<256, 1>(i16, split, p, st);
  if (ana == 256 && nb == 2) return LaunchNsfV<256, 2>(i16, split, p, st);
  if (ana == 256 && nb == 3) return LaunchNsfV<256, 3>(i16, split, p, st);
  return Fail("unsupported (fs, num_bands) combination");
}
// Split mode is sticky: once a stream has seen distinct Analyze / Process signals its two
// histories and magnitude memories differ, so later fused calls on it run the split kernel with
// the Process signal fed to both (identical results to the reference either way).
bool SameAsMemo(const std::vector<Handle*>& hs) {
  if (t_call_is_memo && t_memo.epoch == g_epoch && t_memo.hs.size() == hs.size()) return true;
  return t_memo.epoch == g_epoch && t_memo.hs.size() == hs.size() &&
         memcmp(t_memo.hs.data(), hs.data(), sizeof(Handle*) * hs.size()) == 0;
}
bool NeedSplit(std::vector<Handle*>& hs, bool split_call) {
  if (!split_call && SameAsMemo(hs) && !t_memo.any_split) return false;   // validated list, all fused
  bool split = split_call;
  for (Handle* h : hs) split = split || h->split_mode;
  if (split) {
    bool changed = false;
    for (Handle* h : hs) {
      changed = changed || !h->split_mode;
      h->split_mode = true;
    }
    if (changed) ++g_epoch;
  }
  return split;
}

template <int ANA, int NB>
int LaunchNsxT(const NsxLaunch& p, cudaStream_t st) {
  // CTA size: spread the batch over all SMs, as many lock-stepped warps per CTA as that allows
  // (nsx_kernel.cuh, "CTA shape")
  int dev = 0, sms = 0;
  CU_OK(cudaGetDevice(&dev));
  CU_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int w = (p.n_streams + sms - 1) / sms;
  if (w < kNsxWarpsPerCta) w = kNsxWarpsPerCta;
  if (w > kNsxMaxWarpsPerCta) w = kNsxMaxWarpsPerCta;
  if (const char* e = getenv("NSB200_NSX_WARPS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= kNsxMaxWarpsPerCta) w = v;
  }
  const int grid = (p.n_streams + w - 1) / w;
  NsxLaunch q = p;
  q.prefetch_ahead = sms * w / 2;   // half a resident wave ahead (one CTA per SM), as in the float kernel
  if (const char* e = getenv("NSB200_NSX_AHEAD")) q.prefetch_ahead = atoi(e);
  const size_t smem = sizeof(uint32_t) * (kNsxCtaTableWords + (size_t)w * kNsxWarpWords);
  if (smem > 48 * 1024)   // per device and cheap: set on every large launch
    CU_OK(cudaFuncSetAttribute(nsx_process_kernel<ANA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nsx_process_kernel<ANA, NB><<<grid, w * 32, smem, st>>>(q);
  ++g_launches;
  CU_OK(cudaGetLastError());
  return 0;
}
int LaunchNsx(int ana, int nb, const NsxLaunch& p, cudaStream_t st) {
  if (ana == 128 && nb == 1) return LaunchNsxT<128, 1>(p, st);
  if (ana == 256 && nb == 1) return LaunchNsxT<256, 1>(p, st);
  if (ana == 256 && nb == 2) return LaunchNsxT<256, 2>(p, st);
  if (ana == 256 && nb == 3) return LaunchNsxT<256, 3>(p, st);
  return Fail("unsupported (fs, num_bands) combination");
}

int NumBands(uint32_t fs) { return fs == 32000 ? 2 : (fs == 48000 ? 3 : 1); }

// Host buffers of a host-pointer batch call at 32/48 kHz: the copies become the first and last
// stage of the band pipeline below.
struct HostSub { int first, count, row; };   // batch entries [first, first + count) = staging rows [row, row + count)
struct HostIo {
  const int16_t* in;     // the caller's buffers: entry i at in + i * in_stride
  size_t in_stride;
  int16_t* out;
  size_t out_stride;
  const HostSub* subs;   // which entries this device's staging rows hold
  int nsubs;
};

constexpr int kBandMaxStages = 9;        // copy-in | 7 band / NS stages | copy-out
constexpr int kBandBlockFrames = 100;    // frames per pipelined block (bounds the scratch: ~5.4 KB per stream-frame)

// NS launch over the whole batch, frames [f0, f0 + nf) of PCM laid out with the given strides.
int LaunchNs(DeviceCtx& d, uint32_t magic, int ana, int nb, int n, const int16_t* in, long long in_ss,
             int16_t* out, long long out_ss, long long fstride, long long bstride, int f0, int nf, cudaStream_t s,
             bool split = false, const int16_t* ana_in = nullptr, long long ana_ss = 0, long long ana_fs = 0) {
  if (magic == kMagicF) {
    NsfLaunch p;
    p.state = (float*)d.f_state.base;
    p.hist = (int*)d.f_hist.base;
    p.slots = d.slot_run >= n ? nullptr : d.d_slots;
    p.slot_base = d.slot_run >= n ? d.cached_slots[0] : 0;
    p.tables = d.d_nsf_tables;
    p.in = in + (size_t)f0 * fstride;
    p.out = out + (size_t)f0 * fstride;
    p.in_stream_stride = in_ss;
    p.out_stream_stride = out_ss;
    p.in_frame_stride = p.out_frame_stride = fstride;
    p.in_band_stride = p.out_band_stride = bstride;
    p.n_streams = n;
    p.frames = nf;
    // split without a separate Analyze signal (sticky split mode): band 0 of the input serves both
    p.ana_in = ana_in ? ana_in + (size_t)f0 * ana_fs : in + (size_t)f0 * fstride;
    p.ana_stream_stride = ana_in ? ana_ss : in_ss;
    p.ana_frame_stride = ana_in ? ana_fs : fstride;
    return LaunchNsf(ana, nb, true, split, p, s);
  }
  NsxLaunch p;
  p.state = (uint32_t*)d.x_state.base;
  p.slots = d.slot_run >= n ? nullptr : d.d_slots;
  p.slot_base = d.slot_run >= n ? d.cached_slots[0] : 0;
  p.tables = d.d_nsx_tables;
  p.in = in + (size_t)f0 * fstride;
  p.out = out + (size_t)f0 * fstride;
  p.in_stream_stride = in_ss;
  p.out_stream_stride = out_ss;
  p.in_frame_stride = p.out_frame_stride = fstride;
  p.in_band_stride = p.out_band_stride = bstride;
  p.n_streams = n;
  p.frames = nf;
  return LaunchNsx(ana, nb, p, s);
}

// One block of <= kBandBlockFrames frames through the 32/48 kHz chain.
//
// Every stage kernel walks the frames of its streams serially with filter / resampler state in
// registers, so a stage's duration is set by per-stream latency, not by how many streams it
// carries (measured: 512 and 2048 streams take the same 4.4 ms for 50 frames end to end).  The
// block is therefore cut into chunks of a few frames that flow through one CUDA stream per
// stage: stage s of chunk c waits for stage s-1 of chunk c (event) and for stage s of chunk c-1
// (stream order), and the stages of neighbouring chunks overlap on the otherwise idle SMs.
// With `host`, the H2D and D2H copies are two more stages of the same pipeline.
int RunBandBlock(DeviceCtx& d, uint32_t magic, std::vector<Handle*>& hs, const int16_t* d_in, size_t in_stride,
                 int16_t* d_out, size_t out_stride, int frames, cudaStream_t st, const HostIo* host, bool split) {
  const int n = (int)hs.size();
  const uint32_t fs = hs[0]->fs;
  const int nb = NumBands(fs);
  const int fl = (int)fs / 100;
  const int* d_bslots = d.d_slots + n;

  // band scratch [stream][frame][band][160], split in place of the NS input
  const size_t need = (size_t)n * frames * nb * 160;
  if (need > d.bands_elems) {
    CU_OK(cudaDeviceSynchronize());
    if (d.d_bands) CU_OK(cudaFree(d.d_bands));
    CU_OK(cudaMalloc(&d.d_bands, sizeof(int16_t) * need));
    d.bands_elems = need;
  }
  const size_t need_s = BandScratchElems(nb, n, frames);
  if (need_s > d.band_scratch_elems) {
    CU_OK(cudaDeviceSynchronize());
    if (d.d_band_scratch) CU_OK(cudaFree(d.d_band_scratch));
    CU_OK(cudaMalloc(&d.d_band_scratch, sizeof(int16_t) * need_s));
    d.band_scratch_elems = need_s;
  }
  const long long bands_ss = (long long)frames * nb * 160;

  BandLaunch bl;
  bl.state = (int32_t*)d.b_state.base;
  bl.slots = d_bslots;
  bl.kernel_up = d.d_sinc_up;
  bl.kernel_down = d.d_sinc_down;
  bl.full_in = d_in;
  bl.full_in_stride = (long long)in_stride;
  bl.full_out = d_out;
  bl.full_out_stride = (long long)out_stride;
  bl.bands = d.d_bands;
  bl.bands_stride = bands_ss;
  bl.scratch = d.d_band_scratch;
  bl.n_streams = n;
  bl.frames = frames;

  // 48 kHz merge: 640 -> 480 resampler positions.  Replay the reference's running double per
  // distinct starting value and group the streams whose schedules coincide (streams of different
  // age differ by ~1e-13 in position, which almost never changes a table row or a float weight).
  if (nb == 3) {
    const size_t words = (size_t)frames * 480 * 3;
    std::vector<std::vector<int32_t>> scheds;
    std::vector<std::vector<int>> members;
    std::vector<std::vector<uint8_t>>& regular = d.down_regular;   // per schedule, per frame: fast-kernel eligible
    regular.clear();
    std::vector<std::pair<double, std::pair<int, double>>> seen;   // start -> (schedule, end)
    for (int i = 0; i < n; ++i) {
      const double v0 = hs[i]->down_vsi;
      int gi = -1;
      double vend = 0;
      for (auto& sn : seen)
        if (memcmp(&sn.first, &v0, sizeof(double)) == 0) { gi = sn.second.first; vend = sn.second.second; break; }
      if (gi < 0) {
        std::vector<int32_t> sc(words);
        double v = v0;
        band_down_schedule(&v, frames, sc.data());
        vend = v;
        for (size_t g = 0; g < scheds.size(); ++g)
          if (memcmp(scheds[g].data(), sc.data(), words * sizeof(int32_t)) == 0) { gi = (int)g; break; }
        if (gi < 0) {
          gi = (int)scheds.size();
          regular.emplace_back((size_t)frames);
          for (int f = 0; f < frames; ++f) regular.back()[f] = band_down_frame_regular(sc.data() + (size_t)f * 480 * 3);
          scheds.push_back(std::move(sc));
          members.emplace_back();
        }
        seen.push_back({v0, {gi, vend}});
      }
      members[gi].push_back(i);
      hs[i]->down_vsi = vend;
    }
    const size_t need_w = scheds.size() * words + (scheds.size() > 1 ? (size_t)n : 0);
    if (need_w > d.down_sched_words) {
      CU_OK(cudaDeviceSynchronize());
      if (d.d_down_sched) CU_OK(cudaFree(d.d_down_sched));
      CU_OK(cudaMalloc(&d.d_down_sched, sizeof(int32_t) * need_w));
      d.down_sched_words = need_w;
    }
    int32_t* d_idx = d.d_down_sched + scheds.size() * words;
    size_t idx_off = 0;
    for (size_t g = 0; g < scheds.size(); ++g) {
      // pageable sources: staged by the runtime before the call returns
      CU_OK(cudaMemcpyAsync(d.d_down_sched + g * words, scheds[g].data(), sizeof(int32_t) * words,
                            cudaMemcpyHostToDevice, st));
      BandLaunch::DownGroup dg;
      dg.schedule = d.d_down_sched + g * words;
      dg.count = (int)members[g].size();
      dg.stream_index = nullptr;
      dg.regular = regular[g].data();
      if (scheds.size() > 1) {
        CU_OK(cudaMemcpyAsync(d_idx + idx_off, members[g].data(), sizeof(int) * members[g].size(),
                              cudaMemcpyHostToDevice, st));
        dg.stream_index = reinterpret_cast<const int*>(d_idx + idx_off);
        idx_off += members[g].size();
      }
      bl.down_groups.push_back(dg);
    }
  }

  // chunk plan: short chunks keep every stage busy (measured at 2048 x 48 kHz, ms per 50-frame
  // step: 1-4 frames 2.6, 8 frames 2.8, 25 frames 3.3; unpipelined 4.6)
  int chunk = 4;
  if (const char* e = getenv("NSB200_BAND_CHUNK")) {
    const int v = atoi(e);
    if (v > 0) chunk = v;
  }
  if (chunk > frames) chunk = frames;
  const int nchunks = (frames + chunk - 1) / chunk;
  const int band_stages = BandStages(nb), ns_stage = BandNsStage(nb);
  const int first = host ? 1 : 0;                       // pipeline index of band stage 0
  const int nstages = band_stages + (host ? 2 : 0);
  if (!d.fork_event) {
    CU_OK(cudaEventCreateWithFlags(&d.fork_event, cudaEventDisableTiming));
    // The QMF stages are serial recurrences on a few dozen warps: the pipeline's critical path.  Their streams
    // get the highest priority, so that a chunk's QMF blocks take the next free SM slot ahead of the queued
    // blocks of the wide stages (resamplers, suppressor) of other chunks.
    int prio_lo = 0, prio_hi = 0;
    CU_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (getenv("NSB200_BAND_NOPRIO")) prio_hi = prio_lo;
    for (int k = 0; k < kBandMaxStages; ++k) {
      CU_OK(cudaStreamCreateWithPriority(&d.stage_stream[k], cudaStreamNonBlocking, prio_lo));
      CU_OK(cudaStreamCreateWithPriority(&d.stage_stream_hi[k], cudaStreamNonBlocking, prio_hi));
    }
  }
  while (d.band_events.size() < (size_t)nstages * nchunks) {
    cudaEvent_t e;
    CU_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    d.band_events.push_back(e);
  }
  // NSB200_TRACE=2: timeline of the stages of every chunk on stderr (tuning aid)
  const char* trace_env = getenv("NSB200_TRACE");
  const bool trace = trace_env && atoi(trace_env) >= 2;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t s) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    tev.push_back(e);
  };
  // every stage stream starts after the caller's stream: inputs are ready and the previous
  // block has released the scratch
  CU_OK(cudaEventRecord(d.fork_event, st));
  auto stage_of = [&](int k) -> cudaStream_t {
    const int bk = k - first;   // band stage index: QMF stages are 1, 2, 4, 5 of 7 (48 kHz) or 0, 2 of 3 (32 kHz)
    const bool qmf = bk >= 0 && bk < band_stages && bk != ns_stage && !(nb == 3 && (bk == 0 || bk == 6));
    return qmf ? d.stage_stream_hi[k] : d.stage_stream[k];
  };
  for (int k = 0; k < nstages; ++k) CU_OK(cudaStreamWaitEvent(stage_of(k), d.fork_event, 0));
  mark(stage_of(0));
  const int ana = 256;
  for (int c = 0; c < nchunks; ++c) {
    const int f0 = c * chunk;
    const int nf = frames - f0 < chunk ? frames - f0 : chunk;
    for (int k = 0; k < nstages; ++k) {
      cudaStream_t s = stage_of(k);
      if (k > 0) CU_OK(cudaStreamWaitEvent(s, d.band_events[(size_t)c * nstages + k - 1], 0));
      const int bk = k - first;   // band stage index
      if (bk < 0) {
        for (int u = 0; u < host->nsubs; ++u) {
          const HostSub& sb = host->subs[u];
          CU_OK(cudaMemcpy2DAsync(const_cast<int16_t*>(d_in) + (size_t)sb.row * in_stride + (size_t)f0 * fl, in_stride * sizeof(int16_t),
                                  host->in + (size_t)sb.first * host->in_stride + (size_t)f0 * fl, host->in_stride * sizeof(int16_t),
                                  (size_t)nf * fl * sizeof(int16_t), sb.count, cudaMemcpyHostToDevice, s));
        }
      } else if (bk == band_stages) {
        for (int u = 0; u < host->nsubs; ++u) {
          const HostSub& sb = host->subs[u];
          CU_OK(cudaMemcpy2DAsync(host->out + (size_t)sb.first * host->out_stride + (size_t)f0 * fl, host->out_stride * sizeof(int16_t),
                                  d_out + (size_t)sb.row * out_stride + (size_t)f0 * fl, out_stride * sizeof(int16_t),
                                  (size_t)nf * fl * sizeof(int16_t), sb.count, cudaMemcpyDeviceToHost, s));
        }
      } else if (bk == ns_stage) {
        if (LaunchNs(d, magic, ana, nb, n, d.d_bands, bands_ss, d.d_bands, bands_ss, nb * 160, 160, f0, nf, s, split) != 0)
          return -1;
      } else {
        uint64_t launched = 0;
        if (LaunchBandStage(nb, bl, bk, f0, nf, s, &launched) != 0) return Fail("band stage launch failed");
        g_launches += launched;
      }
      CU_OK(cudaEventRecord(d.band_events[(size_t)c * nstages + k], s));
      mark(s);
    }
  }
  // join: the last stage of the last chunk follows everything else of the block
  CU_OK(cudaStreamWaitEvent(st, d.band_events[(size_t)(nchunks - 1) * nstages + nstages - 1], 0));
  if (trace) {
    CU_OK(cudaDeviceSynchronize());
    for (int c = 0; c < nchunks; ++c) {
      fprintf(stderr, "chunk %2d ends:", c);
      for (int k = 0; k < nstages; ++k) {
        float t;
        cudaEventElapsedTime(&t, tev[0], tev[1 + (size_t)c * nstages + k]);
        fprintf(stderr, " %.3f", t);
      }
      fprintf(stderr, " ms\n");
    }
    for (auto e : tev) cudaEventDestroy(e);
  }
  return 0;
}

// Enqueues the suppressor (with band split / merge at 32/48 kHz) for the streams of one device.
// d_in / d_out: full-band int16 PCM in device memory; with `host` (32/48 kHz only) they are
// staging buffers the pipeline fills from / drains to the host buffers itself.
// d_ana (float NS at 8/16 kHz only): separate signal for Analyze, [stream][frame][fs/100] int16.
int RunDevice(DeviceCtx& d, uint32_t magic, std::vector<Handle*>& hs, const int16_t* d_in,
              size_t in_stride, int16_t* d_out, size_t out_stride, int frames, cudaStream_t st,
              const HostIo* host = nullptr, const int16_t* d_ana = nullptr, size_t ana_stride = 0) {
  const int n = (int)hs.size();
  const uint32_t fs = hs[0]->fs;
  const int nb = NumBands(fs);
  const int fl = (int)fs / 100;
  // slot lists: [0,n) = NS slots, [n,2n) = band slots (remembered for the validated list)
  const bool memo = SameAsMemo(hs);
  std::vector<int> built;
  if (!memo || t_memo.slots.empty()) {
    built.resize(nb > 1 ? 2 * (size_t)n : (size_t)n);
    for (int i = 0; i < n; ++i) {
      built[i] = hs[i]->slot;
      if (nb > 1) built[n + i] = BandSlot(hs[i]);
    }
    if (memo) t_memo.slots = built;
  }
  const std::vector<int>& all = (memo && !t_memo.slots.empty()) ? t_memo.slots : built;
  if (UploadSlots(d, all, st) != 0) return -1;
  if (d_ana && (magic != kMagicF || nb != 1))
    return Fail("a separate Analyze signal is supported for the float suppressor on full-band PCM at 8/16 kHz "
                "(use WebRtcNs_AnalyzeProcessBatchBandsF32 on band frames at 32/48 kHz)");
  const bool split = magic == kMagicF && NeedSplit(hs, d_ana != nullptr);
  if (nb == 1)
    return LaunchNs(d, magic, fs == 8000 ? 128 : 256, 1, n, d_in, (long long)in_stride, d_out, (long long)out_stride,
                    fl, 0, 0, frames, st, split, d_ana, (long long)ana_stride, fl);
  if ((in_stride | out_stride) & 7) return Fail("strides must be multiples of 8 samples at 32/48 kHz");
  for (int f0 = 0; f0 < frames; f0 += kBandBlockFrames) {
    const int nf = frames - f0 < kBandBlockFrames ? frames - f0 : kBandBlockFrames;
    HostIo h;
    if (host) {
      h = *host;
      h.in += (size_t)f0 * fl;
      h.out += (size_t)f0 * fl;
    }
    if (RunBandBlock(d, magic, hs, d_in + (size_t)f0 * fl, in_stride, d_out + (size_t)f0 * fl, out_stride, nf, st,
                     host ? &h : nullptr, split) != 0)
      return -1;
  }
  return 0;
}

// Validates a batch call.  in / out (may be NULL for calls that carry no PCM): the kernels move PCM as 32-bit
// words (16-byte vectors at 32/48 kHz), so the pointers must be aligned to that; strides likewise.  A handle
// listed twice would put two warps on one state slab.  With one stream the stride is never used to step.
int CheckBatch(void* const* hv, int n, uint32_t magic, size_t* in_stride, size_t* out_stride,
               int frames, std::vector<Handle*>* hs, const void* in = nullptr, const void* out = nullptr) {
  if (!hv || n <= 0) return Fail("no handles");
  if (frames < 0) return Fail("negative frame count");
  if ((*in_stride | *out_stride) & 1) return Fail("strides must be even");
  if (t_memo.epoch == g_epoch && t_memo.magic == magic && t_memo.hv.size() == (size_t)n &&
      memcmp(t_memo.hv.data(), hv, sizeof(void*) * (size_t)n) == 0) {
    *hs = t_memo.hs;
  } else {
    hs->resize(n);
    bool any_split = false;
    int single_dev = -1;
    const uint64_t stamp = ++g_call_stamp;
    for (int i = 0; i < n; ++i) {
      Handle* h = AsHandle(hv[i], magic);
      if (!h) return Fail("bad handle in batch");
      if (!h->init_flag) return Fail("handle not initialised");
      if (h->fs != static_cast<Handle*>(hv[0])->fs) return Fail("mixed sample rates in one batch");
      if (h->seen_stamp == stamp) return Fail("handle listed twice in one batch");
      h->seen_stamp = stamp;
      any_split = any_split || h->split_mode;
      single_dev = i == 0 ? h->dev : (single_dev == h->dev ? single_dev : -1);
      (*hs)[i] = h;
    }
    t_memo.epoch = g_epoch;
    t_memo.magic = magic;
    t_memo.hv.assign(hv, hv + n);
    t_memo.hs = *hs;
    t_memo.slots.clear();
    t_memo.any_split = any_split;
    t_memo.single_dev = single_dev;
  }
  const uint32_t fs = (*hs)[0]->fs;
  const size_t need = (size_t)frames * (fs / 100);
  if (n == 1) {   // one row: any stride will do, and a 2-D copy must not see a pitch below its width
    if (*in_stride < need) *in_stride = need;
    if (*out_stride < need) *out_stride = need;
  } else if (*in_stride < need || *out_stride < need) {
    return Fail("stride shorter than the frames of one stream");
  }
  const uintptr_t align = fs > 16000 ? 15u : 3u;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & align)
    return Fail(fs > 16000 ? "PCM pointers must be 16-byte aligned at 32/48 kHz" : "PCM pointers must be 4-byte aligned");
  t_call_is_memo = true;
  return 0;
}

// A launch a *Device entry point put on the caller's stream: the library's own stream (and with it every
// later call that synchronises or enqueues there: Free, Init, set_policy, the getter, host-pointer batches)
// is ordered behind it.
int OrderAfterUserStream(DeviceCtx& d, cudaStream_t st) {
  if (st == d.stream) return 0;
  if (!d.user_done) CU_OK(cudaEventCreateWithFlags(&d.user_done, cudaEventDisableTiming));
  CU_OK(cudaEventRecord(d.user_done, st));
  CU_OK(cudaStreamWaitEvent(d.stream, d.user_done, 0));
  return 0;
}

int BatchDevice(void* const* hv, int n, uint32_t magic, const int16_t* in, size_t in_stride,
                int16_t* out, size_t out_stride, int frames, void* stream) {
  UseLock lk;
  static thread_local std::vector<Handle*> hs;   // keeps its capacity: no 256 KB allocation per tick
  if (CheckBatch(hv, n, magic, &in_stride, &out_stride, frames, &hs, in, out) != 0) return -1;
  if (frames == 0) return 0;
  if (t_memo.single_dev < 0) return Fail("device batch spans several GPUs");
  lk.Device(hs[0]->dev);
  DeviceCtx* d;
  if (DeviceReady(hs[0]->dev, &d) != 0) return -1;
  cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
  if (RunDevice(*d, magic, hs, in, in_stride, out, out_stride, frames, st) != 0) return -1;
  return OrderAfterUserStream(*d, st);
}

// Host-pointer batch: the handles are bucketed by device (any order, any interleaving of devices in the
// list); per device copy in -> run -> copy out.  Frames are cut into chunks so that the H2D copy of chunk
// c+1 and the D2H copy of chunk c-1 overlap the kernel of chunk c (three streams, events).
// A bucket's streams sit in its staging buffers in batch order; runs of consecutive batch entries move with
// one 2-D copy each.
struct HostBucket {
  int dev = -1;
  std::vector<HostSub> subs;
  std::vector<Handle*> hs;
};

int BatchHost(void* const* hv, int n, uint32_t magic, const int16_t* in, size_t in_stride,
              int16_t* out, size_t out_stride, int frames, uint64_t* ticket = nullptr) {
  // ticket: asynchronous call -- everything is enqueued behind what earlier asynchronous calls
  // left in the three queues and the call returns without waiting (0 = nothing to wait for)
  UseLock lk;
  if (ticket) *ticket = 0;
  std::vector<Handle*> hs;
  if (CheckBatch(hv, n, magic, &in_stride, &out_stride, frames, &hs, in, out) != 0) return -1;
  if (frames == 0) return 0;
  const int fl = (int)hs[0]->fs / 100;
  std::vector<HostBucket> buckets;
  {
    std::vector<int> of_dev(g_devs.size(), -1);
    for (int i = 0; i < n; ++i) {
      const int dv = hs[i]->dev;
      if (of_dev[dv] < 0) {
        of_dev[dv] = (int)buckets.size();
        buckets.emplace_back();
        buckets.back().dev = dv;
      }
      HostBucket& b = buckets[of_dev[dv]];
      if (!b.subs.empty() && b.subs.back().first + b.subs.back().count == i) b.subs.back().count++;
      else b.subs.push_back(HostSub{i, 1, (int)b.hs.size()});
      b.hs.push_back(hs[i]);
    }
    std::sort(buckets.begin(), buckets.end(), [](const HostBucket& x, const HostBucket& y) { return x.dev < y.dev; });
  }
  const bool bands = NumBands(hs[0]->fs) > 1;
  // (asynchronous calls at 32/48 kHz: that path blocks, after whatever is still in flight)
  for (const HostBucket& b : buckets) lk.Device(b.dev, ticket == nullptr || bands);
  if (bands) {
    // 32/48 kHz: the copies are the first and last stage of the band pipeline (RunBandBlock);
    // staging holds one block of frames per direction
    const int block = frames < kBandBlockFrames ? frames : kBandBlockFrames;
    for (const HostBucket& b : buckets) {
      DeviceCtx* d;
      if (DeviceReady(b.dev, &d) != 0) return -1;
      const size_t need = b.hs.size() * (size_t)block * fl;
      if (need > d->stage_elems) {
        CU_OK(cudaDeviceSynchronize());
        if (d->d_in) { CU_OK(cudaFree(d->d_in)); CU_OK(cudaFree(d->d_out)); }
        CU_OK(cudaMalloc(&d->d_in, sizeof(int16_t) * need));
        CU_OK(cudaMalloc(&d->d_out, sizeof(int16_t) * need));
        d->stage_elems = need;
        d->pipe_per = 0;
      }
    }
    for (int f0 = 0; f0 < frames; f0 += block) {
      const int nf = frames - f0 < block ? frames - f0 : block;
      for (HostBucket& b : buckets) {
        DeviceCtx* d = &g_devs[b.dev];
        CU_OK(cudaSetDevice(b.dev));
        HostIo hio = {in + (size_t)f0 * fl, in_stride, out + (size_t)f0 * fl, out_stride, b.subs.data(), (int)b.subs.size()};
        const size_t per = (size_t)block * fl;
        if (RunDevice(*d, magic, b.hs, d->d_in, per, d->d_out, per, nf, d->stream, &hio) != 0) return -1;
      }
    }
    for (const HostBucket& b : buckets) {
      CU_OK(cudaSetDevice(b.dev));
      CU_OK(cudaStreamSynchronize(g_devs[b.dev].stream));
    }
    return 0;
  }
  // chunking: the copies bound the call (PCIe, ~50 GB/s each way, full duplex) as soon as the
  // kernel keeps up, so the pipeline wants many small chunks -- fill and drain cost one chunk of
  // copy-in, kernel and copy-out each -- but chunks large enough that DMA set-up, launch overhead
  // and the kernel's per-launch state round trip stay small: ~12 MB per direction per chunk.
  int chunk = frames;
  {
    const double bytes_per_frame = (double)n * fl * sizeof(int16_t);
    // ~12 MB per chunk, but at least four chunks when the call is big enough for 1.5 MB chunks
    // (asynchronous calls queue behind each other, so fill and drain are paid once per stream of
    // calls, not per call: ~32 MB chunks measured best, tools/e2e_sweep.sh)
    double target = ticket ? 32.0e6 : 12.0e6;
    if (!ticket && bytes_per_frame * frames / 4.0 < target) target = bytes_per_frame * frames / 4.0;
    if (target < 1.5e6) target = 1.5e6;
    int c = (int)(target / bytes_per_frame + 0.5);
    if (c < 1) c = 1;
    if (c > 250) c = 250;
    if (const char* e = getenv("NSB200_CHUNK_FRAMES")) {
      const int v = atoi(e);
      if (v > 0) c = v;
    }
    if (c < chunk) chunk = c;
  }
  // chunk plan: half-size first and last chunks shorten the pipeline's fill (first copy-in) and
  // drain (last kernel + last copy-out), which nothing overlaps
  std::vector<int> starts;   // first frame of each chunk, plus the end
  {
    const int edge = !ticket && chunk >= 4 && frames >= 3 * chunk ? chunk / 2 : chunk;
    int f0 = 0;
    starts.push_back(0);
    f0 += edge < frames ? edge : frames;
    while (frames - f0 > chunk + edge) { starts.push_back(f0); f0 += chunk; }
    if (f0 < frames) {
      starts.push_back(f0);
      if (frames - f0 > chunk) starts.push_back(frames - edge);
    }
    starts.push_back(frames);
  }
  const int nchunks = (int)starts.size() - 1;
  int kBuf = 3;   // staging buffers per direction
  if (const char* e = getenv("NSB200_STAGE_BUFS")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 8) kBuf = v;
  }
  for (const HostBucket& b : buckets) {
    DeviceCtx* d;
    if (DeviceReady(b.dev, &d) != 0) return -1;
    const int count = (int)b.hs.size();
    const size_t per = (size_t)chunk * fl;           // samples per stream per chunk
    const size_t need = kBuf * (size_t)count * per;
    if (need > d->stage_elems) {
      CU_OK(cudaDeviceSynchronize());
      if (d->d_in) { CU_OK(cudaFree(d->d_in)); CU_OK(cudaFree(d->d_out)); }
      CU_OK(cudaMalloc(&d->d_in, sizeof(int16_t) * need));
      CU_OK(cudaMalloc(&d->d_out, sizeof(int16_t) * need));
      d->stage_elems = need;
      d->pipe_per = 0;
    }
    if (d->pipe_per != per || d->pipe_count != count || d->pipe_bufs != kBuf) {
      // the buffers are cut differently from the pipeline still in flight: let it run dry
      CU_OK(cudaStreamSynchronize(d->copy_in));
      CU_OK(cudaStreamSynchronize(d->stream));
      CU_OK(cudaStreamSynchronize(d->copy_out));
      d->pipe_per = per;
      d->pipe_count = count;
      d->pipe_bufs = kBuf;
      for (int k = 0; k < kBuf; ++k) d->pipe_used[k] = false;
    }
    for (int k = 0; k < kBuf; ++k)
      if (!d->pipe_kdone[k]) {
        CU_OK(cudaEventCreateWithFlags(&d->pipe_kdone[k], cudaEventDisableTiming));
        CU_OK(cudaEventCreateWithFlags(&d->pipe_drained[k], cudaEventDisableTiming));
      }
    const size_t nev = 3 * (size_t)nchunks;
    while (d->events.size() < nev) {
      cudaEvent_t e;
      CU_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      d->events.push_back(e);
    }
  }
  // One device at a time issues its whole pipeline asynchronously; devices run
  // concurrently because nothing below blocks the host until the final syncs.
  for (HostBucket& b : buckets) {
    DeviceCtx* d = &g_devs[b.dev];
    CU_OK(cudaSetDevice(b.dev));
    const int count = (int)b.hs.size();
    const size_t per = (size_t)chunk * fl;
    std::vector<cudaEvent_t>& ev = d->events;
    // NSB200_TRACE=1: per-chunk timeline of the three queues on stderr (tuning aid)
    const bool trace = getenv("NSB200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
      if (!trace) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, st);
      tev.push_back(e);
    };
    for (int c = 0; c < nchunks; ++c) {
      const int f0 = starts[c];
      const int nf = starts[c + 1] - f0;
      const int k = (int)((d->pipe_seq + (unsigned long long)c) % (unsigned long long)kBuf);
      int16_t* din = d->d_in + (size_t)k * count * per;
      int16_t* dout = d->d_out + (size_t)k * count * per;
      // buffer k was last read by the kernel three chunks back and last drained by its copy-out
      // (of this call or, for asynchronous calls, of the one before)
      if (d->pipe_used[k]) {
        CU_OK(cudaStreamWaitEvent(d->copy_in, d->pipe_kdone[k], 0));
        CU_OK(cudaStreamWaitEvent(d->stream, d->pipe_drained[k], 0));
      }
      mark(d->copy_in);
      for (const HostSub& sb : b.subs)
        CU_OK(cudaMemcpy2DAsync(din + (size_t)sb.row * per, per * sizeof(int16_t),
                                in + (size_t)sb.first * in_stride + (size_t)f0 * fl, in_stride * sizeof(int16_t),
                                (size_t)nf * fl * sizeof(int16_t), sb.count, cudaMemcpyHostToDevice, d->copy_in));
      mark(d->copy_in);
      CU_OK(cudaEventRecord(ev[3 * c + 0], d->copy_in));
      CU_OK(cudaStreamWaitEvent(d->stream, ev[3 * c + 0], 0));
      mark(d->stream);
      if (RunDevice(*d, magic, b.hs, din, per, dout, per, nf, d->stream) != 0) return -1;
      mark(d->stream);
      CU_OK(cudaEventRecord(d->pipe_kdone[k], d->stream));
      CU_OK(cudaStreamWaitEvent(d->copy_out, d->pipe_kdone[k], 0));
      mark(d->copy_out);
      for (const HostSub& sb : b.subs)
        CU_OK(cudaMemcpy2DAsync(out + (size_t)sb.first * out_stride + (size_t)f0 * fl, out_stride * sizeof(int16_t),
                                dout + (size_t)sb.row * per, per * sizeof(int16_t),
                                (size_t)nf * fl * sizeof(int16_t), sb.count, cudaMemcpyDeviceToHost, d->copy_out));
      mark(d->copy_out);
      CU_OK(cudaEventRecord(d->pipe_drained[k], d->copy_out));
      d->pipe_used[k] = true;
    }
    d->pipe_seq += (unsigned long long)nchunks;
    if (trace) {
      CU_OK(cudaDeviceSynchronize());
      for (int c = 0; c < nchunks; ++c) {
        float t[6];
        for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], tev[0], tev[6 * c + k]);
        fprintf(stderr, "chunk %2d  h2d %.3f-%.3f  kernel %.3f-%.3f  d2h %.3f-%.3f ms\n", c, t[0], t[1], t[2], t[3], t[4], t[5]);
      }
      for (auto e : tev) cudaEventDestroy(e);
    }
  }
  if (ticket) {
    // completion = the last copy-out of every device (the copy-out queue is in order)
    PendingBatch pb;
    for (const HostBucket& b : buckets) {
      DeviceCtx* d = &g_devs[b.dev];
      CU_OK(cudaSetDevice(b.dev));
      cudaEvent_t e;
      CU_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CU_OK(cudaEventRecord(e, d->copy_out));
      pb.done.push_back(std::make_pair(b.dev, e));
    }
    std::lock_guard<std::mutex> g(g_pend_mu);
    pb.ticket = g_next_ticket++;
    *ticket = pb.ticket;
    g_pending.push_back(pb);
    return 0;
  }
  for (const HostBucket& b : buckets) {
    DeviceCtx* d = &g_devs[b.dev];
    CU_OK(cudaSetDevice(b.dev));
    CU_OK(cudaStreamSynchronize(d->copy_out));
    CU_OK(cudaStreamSynchronize(d->stream));
  }
  return 0;
}

int WaitBatch(uint64_t ticket) {
  // tickets complete in issue order per device, but callers may wait in any order
  PendingBatch mine;
  mine.ticket = 0;
  {
    std::lock_guard<std::mutex> g(g_pend_mu);
    for (size_t i = 0; i < g_pending.size(); ++i)
      if (g_pending[i].ticket == ticket) {
        mine = std::move(g_pending[i]);
        g_pending.erase(g_pending.begin() + (long)i);
        break;
      }
  }
  FinishPending(mine);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return Fail(std::string("asynchronous batch failed: ") + cudaGetErrorString(e));
  return 0;
}

// Single-stream float call = batch of one over float band frames.
// ana (optional): [stream][frame][frame_len] band-0 frames for Analyze, stride ana_ss floats.
// phase (nsf_kernel.cuh): 0 = whole frames; 1 = the Analyze half alone (only `ana` is read, nothing written
// but state); 2 = the Process half alone on `in`, using what the Analyze call left in the state.
int ProcessBandsF32(void* const* hv, int n, int nb, const float* in, size_t in_ss, float* out,
                    size_t out_ss, int frames, const float* ana = nullptr, size_t ana_ss = 0, int phase = 0) {
  UseLock lk;
  if (!hv || n <= 0) return Fail("no handles");
  if (phase == 1 ? !ana : (!in || !out)) return Fail("NULL frame pointer");
  if (nb < 1 || nb > 3) return Fail("num_bands out of range");
  std::vector<Handle*> hs(n);
  const uint64_t stamp = ++g_call_stamp;
  for (int i = 0; i < n; ++i) {
    Handle* h = AsHandle(hv[i], kMagicF);
    if (!h) return Fail("bad handle in batch");
    if (!h->init_flag) return Fail("handle not initialised");
    if (h->fs != static_cast<Handle*>(hv[0])->fs || h->dev != static_cast<Handle*>(hv[0])->dev)
      return Fail("mixed sample rates or devices in one batch");
    if (h->seen_stamp == stamp) return Fail("handle listed twice in one batch");
    h->seen_stamp = stamp;
    hs[i] = h;
  }
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ana)) & 3u)
    return Fail("float frames must be 4-byte aligned");
  lk.Device(hs[0]->dev);
  if (frames <= 0) return frames == 0 ? 0 : Fail("negative frame count");
  const uint32_t fs = hs[0]->fs;
  if (fs == 8000 && nb != 1) return Fail("8 kHz has a single band");
  const int fl = fs == 8000 ? 80 : 160;
  DeviceCtx* d;
  if (DeviceReady(hs[0]->dev, &d) != 0) return -1;
  const size_t per = (size_t)frames * nb * fl, per_a = (size_t)frames * fl;
  if (n > 1 && (in_ss < per || out_ss < per || (ana && ana_ss < per_a))) return Fail("stride shorter than the frames of one stream");
  if (n == 1) {   // one row: the stride is never used to step; a 2-D copy must not see a pitch below its width
    in_ss = out_ss = per;
    ana_ss = per_a;
  }
  const bool split = NeedSplit(hs, ana != nullptr || phase != 0);
  float *din = nullptr, *dout = nullptr, *dana = nullptr;
  // single-frame calls on one stream (the reference's own calling pattern) reuse a small per-device scratch
  const bool small = n == 1 && frames == 1;
  if (small && !d->d_single) CU_OK(cudaMalloc(&d->d_single, sizeof(float) * (3 * 160 * 2 + 160)));
  if (small) {
    din = d->d_single;
    dout = din + 3 * 160;
    dana = ana ? dout + 3 * 160 : nullptr;
  } else {
    CU_OK(cudaMalloc(&din, sizeof(float) * per * n));
    CU_OK(cudaMalloc(&dout, sizeof(float) * per * n));
    if (ana) CU_OK(cudaMalloc(&dana, sizeof(float) * per_a * n));
  }
  int rc = 0;
  do {
    std::vector<int> slots(n);
    for (int i = 0; i < n; ++i) slots[i] = hs[i]->slot;
    if ((rc = UploadSlots(*d, slots, d->stream)) != 0) break;
    cudaError_t e = cudaSuccess;
    if (phase != 1)
      e = cudaMemcpy2DAsync(din, per * sizeof(float), in, in_ss * sizeof(float), per * sizeof(float), n,
                            cudaMemcpyHostToDevice, d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
    if (ana) {
      e = cudaMemcpy2DAsync(dana, per_a * sizeof(float), ana, ana_ss * sizeof(float), per_a * sizeof(float), n,
                            cudaMemcpyHostToDevice, d->stream);
      if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
    }
    NsfLaunch p;
    p.state = (float*)d->f_state.base;
    p.hist = (int*)d->f_hist.base;
    p.slots = d->d_slots;
    p.tables = d->d_nsf_tables;
    p.in = din;
    p.out = dout;
    p.in_stream_stride = p.out_stream_stride = (long long)per;
    p.in_frame_stride = p.out_frame_stride = (long long)nb * fl;
    p.in_band_stride = p.out_band_stride = fl;
    p.n_streams = n;
    p.frames = frames;
    p.ana_in = ana ? dana : din;
    p.ana_stream_stride = ana ? (long long)per_a : (long long)per;
    p.ana_frame_stride = ana ? (long long)fl : (long long)nb * fl;
    p.phase = phase;
    if ((rc = LaunchNsf(fs == 8000 ? 128 : 256, nb, false, split, p, d->stream)) != 0) break;
    if (phase != 1)
      e = cudaMemcpy2DAsync(out, out_ss * sizeof(float), dout, per * sizeof(float), per * sizeof(float), n,
                            cudaMemcpyDeviceToHost, d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
    e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
  } while (0);
  if (!small) {
    cudaFree(din);
    cudaFree(dout);
    if (dana) cudaFree(dana);
  }
  return rc;
}

// Full-band int16 PCM with a separate Analyze signal, host pointers (8/16 kHz, float NS).
int SplitBatchHost(void* const* hv, int n, const int16_t* ana, size_t ana_stride, const int16_t* in, size_t in_stride,
                   int16_t* out, size_t out_stride, int frames) {
  UseLock lk;
  std::vector<Handle*> hs;
  if (!ana) return Fail("NULL Analyze signal");
  if (CheckBatch(hv, n, kMagicF, &in_stride, &out_stride, frames, &hs, in, out) != 0) return -1;
  if (frames == 0) return 0;
  const int fl = (int)hs[0]->fs / 100;
  const size_t per = (size_t)frames * fl;
  if ((ana_stride & 1) || (n > 1 && ana_stride < per) || (reinterpret_cast<uintptr_t>(ana) & 3u)) return Fail("bad Analyze stride or alignment");
  if (n == 1 && ana_stride < per) ana_stride = per;
  if (t_memo.single_dev < 0) return Fail("split batch spans several GPUs");
  lk.Device(hs[0]->dev);
  DeviceCtx* d;
  if (DeviceReady(hs[0]->dev, &d) != 0) return -1;
  int16_t* buf = nullptr;
  CU_OK(cudaMalloc(&buf, sizeof(int16_t) * per * n * 3));
  int16_t *dana = buf, *din = buf + per * n, *dout = buf + 2 * per * n;
  int rc = 0;
  do {
    cudaError_t e = cudaMemcpy2DAsync(dana, per * 2, ana, ana_stride * 2, per * 2, n, cudaMemcpyHostToDevice, d->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(din, per * 2, in, in_stride * 2, per * 2, n, cudaMemcpyHostToDevice, d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
    if ((rc = RunDevice(*d, kMagicF, hs, din, per, dout, per, frames, d->stream, nullptr, dana, per)) != 0) break;
    e = cudaMemcpy2DAsync(out, out_stride * 2, dout, per * 2, per * 2, n, cudaMemcpyDeviceToHost, d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
  } while (0);
  cudaFree(buf);
  return rc;
}
int SplitBatchDevice(void* const* hv, int n, const int16_t* ana, size_t ana_stride, const int16_t* in, size_t in_stride,
                     int16_t* out, size_t out_stride, int frames, void* stream) {
  UseLock lk;
  std::vector<Handle*> hs;
  if (!ana) return Fail("NULL Analyze signal");
  if (CheckBatch(hv, n, kMagicF, &in_stride, &out_stride, frames, &hs, in, out) != 0) return -1;
  if (frames == 0) return 0;
  if ((ana_stride & 1) || (n > 1 && ana_stride < (size_t)frames * (hs[0]->fs / 100)) || (reinterpret_cast<uintptr_t>(ana) & 3u))
    return Fail("bad Analyze stride or alignment");
  if (t_memo.single_dev < 0) return Fail("device batch spans several GPUs");
  lk.Device(hs[0]->dev);
  DeviceCtx* d;
  if (DeviceReady(hs[0]->dev, &d) != 0) return -1;
  cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
  if (RunDevice(*d, kMagicF, hs, in, in_stride, out, out_stride, frames, st, nullptr, ana, ana_stride) != 0) return -1;
  return OrderAfterUserStream(*d, st);
}

// ---- interleaved multi-channel front end (APM_NS::processCaptureStream, libapm/src/apm_ns.cpp:47-132)
// FloatToS16 / S16ToFloat of common_audio/include/audio_util.h:27-39.
__device__ __forceinline__ int float_to_s16(float v) {
  if (v > 0.f) return v >= 1.f ? 32767 : (int)(int16_t)(v * 32767.f + 0.5f);
  return v <= -1.f ? -32768 : (int)(int16_t)(-v * -32768.f - 0.5f);
}
__device__ __forceinline__ float s16_to_float(int v) {
  const float kMaxInv = 1.f / 32767.f, kMinInv = 1.f / -32768.f;
  return (float)v * (v > 0 ? kMaxInv : -kMinInv);
}
// interleaved [sample][channel] -> planar [channel][samples] int16
template <typename T>
__global__ void deinterleave_kernel(const T* in, int16_t* out, int channels, int samples) {
  const size_t total = (size_t)channels * samples;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const size_t n = i / channels;
    int v;
    if (sizeof(T) == 2) v = (int)reinterpret_cast<const int16_t*>(in)[i];
    else v = float_to_s16(reinterpret_cast<const float*>(in)[i]);
    out[(size_t)c * samples + n] = (int16_t)v;
  }
}
template <typename T>
__global__ void interleave_kernel(const int16_t* in, T* out, int channels, int samples) {
  const size_t total = (size_t)channels * samples;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const size_t n = i / channels;
    const int v = in[(size_t)c * samples + n];
    if (sizeof(T) == 2) reinterpret_cast<int16_t*>(out)[i] = (int16_t)v;
    else reinterpret_cast<float*>(out)[i] = s16_to_float(v);
  }
}

// Device self-test of the arithmetic shortcuts over pseudo-random operands in the kernels'
// ranges: nsb_logf / nsb_sqrtf_p1 / round_s16 / fx_sqrt_floor() must equal their definitions
// (logf, sqrtf + 1, FloatS16ToS16, spl_sqrt_floor.c:55) bit for bit; fdiv() is compared with IEEE
// division and classified: equal, one ulp off, worse.  out[0] = hard mismatches (incl. divisions
// more than one ulp off), out[1] = divisions one ulp off, out[2] = divisions checked, out[3] = nsb_log_rn()
// results that are not (float)log((double)x) (both are the correctly rounded float except within ~2^-41 /
// 2^-52 of a rounding boundary: a few per million at most), out[4] = logarithms checked, out[5] = nsb_exp_rn()
// results that are not (float)exp((double)x), out[6] = exponentials checked, out[7] = sigmoid maps
// 0.5f * (nsb_tanh_rn(x) + 1.f) that differ from the one built on the library's tanh (checked as often as out[6]).
__device__ __forceinline__ void selftest_div(float got, float want, unsigned long long& hard,
                                             unsigned long long& ulp1, unsigned long long& ndiv) {
  ++ndiv;
  const int d = __float_as_int(got) - __float_as_int(want);
  if (d == 1 || d == -1) ++ulp1;
  else if (d != 0) ++hard;
}
__global__ void selftest_kernel(unsigned long long n, unsigned seed, unsigned long long* out) {
  unsigned long long mism = 0, ulp1 = 0, ndiv = 0, lrn = 0, nlog = 0, ern = 0, nexp = 0, trn = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t h1 = pcm_mix32(seed + (uint32_t)i * 2u + (uint32_t)(i >> 31));
    const uint32_t h2 = pcm_mix32(h1 ^ 0x9E3779B9u);
    // floats with exponents in [2^-20, 2^40): sign from the hash
    const float a = __uint_as_float((h1 & 0x807fffffu) | (((h1 >> 23) % 60u + 107u) << 23));
    const float b = __uint_as_float((h2 & 0x007fffffu) | (((h2 >> 23) % 60u + 107u) << 23));
    selftest_div(fdiv(a, b), __fdiv_rn(a, b), mism, ulp1, ndiv);
    // divisions by compile-time constants with RN(1/b) as the starting reciprocal
    selftest_div(NSB_FDIV_C(a, 129.f), __fdiv_rn(a, 129.f), mism, ulp1, ndiv);
    selftest_div(NSB_FDIV_C(a, 65.f), __fdiv_rn(a, 65.f), mism, ulp1, ndiv);
    selftest_div(NSB_FDIV_C(a, 0.1f), __fdiv_rn(a, 0.1f), mism, ulp1, ndiv);
    selftest_div(NSB_FDIV_C(a, 0.05f), __fdiv_rn(a, 0.05f), mism, ulp1, ndiv);
    {
      // shared reciprocal: counter + 1 in 1..201 under quantile-tracker numerators
      const float cb = (float)(h2 % 201u + 1u);
      selftest_div(fdiv_r(a, cb, frcp_nr(cb)), __fdiv_rn(a, cb), mism, ulp1, ndiv);
      // round_s16 against the reference's branches
      const float v = __uint_as_float((h1 & 0x807fffffu) | (((h1 >> 23) % 24u + 120u) << 23));   // |v| in [2^-7, 2^17)
      const int want = v > 0.f ? (v >= 32766.5f ? 32767 : (int)(v + 0.5f)) : (v <= -32767.5f ? -32768 : (int)(v - 0.5f));
      if (round_s16(v) != want) ++mism;
    }
    {
      // nsb_logf == logf on [1, 2^40) (spectral magnitudes + 1, 1 + 2 snrPrior)
      const float x = __uint_as_float((h1 & 0x007fffffu) | (((h2 >> 9) % 40u + 127u) << 23));
      if (__float_as_uint(nsb_logf(x)) != __float_as_uint(logf(x))) ++mism;
      const float x1 = 1.f + __uint_as_float((h2 & 0x007fffffu) | (((h1 >> 9) % 30u + 97u) << 23));  // just above 1
      if (__float_as_uint(nsb_logf(x1)) != __float_as_uint(logf(x1))) ++mism;
      // nsb_log_rn against the double-precision library logarithm rounded to float, and never more than an ulp off
      const float xm = __uint_as_float((h2 & 0x007fffffu) | (((h1 >> 11) % 26u + 127u) << 23));   // magnitudes + 1: [1, 2^26)
      for (int t = 0; t < 2; ++t) {
        const float xx = t ? x1 : xm;
        const float got = nsb_log_rn(xx), want = (float)log((double)xx);
        const int d = __float_as_int(got) - __float_as_int(want);
        ++nlog;
        if (d != 0) ++lrn;
        if (d > 1 || d < -1) ++mism;
      }
      // nsb_exp_rn on [-40, 40] (-logLrt, lquantile, the flatness mean) and beyond both ends of the float range
      const float xe = ((float)(h1 >> 8) * (1.f / 16777216.f) - 0.5f) * ((h2 & 7u) == 0 ? 260.f : 80.f);
      {
        const float got = nsb_exp_rn(xe), want = (float)exp((double)xe);
        const int d = __float_as_int(got) - __float_as_int(want);
        ++nexp;
        if (d != 0) ++ern;
        if (d > 1 || d < -1) ++mism;
      }
      // the sigmoid map of SpeechNoiseProb built on nsb_tanh_rn: arguments from 1e-6 to +-50
      const float xt = __uint_as_float((h2 & 0x807fffffu) | (((h1 >> 5) % 26u + 107u) << 23));
      if (__float_as_uint(0.5f * (nsb_tanh_rn(xt) + 1.f)) != __float_as_uint(0.5f * ((float)tanh((double)xt) + 1.f))) ++trn;
    }
    {
      // nsb_sqrtf_p1 == sqrtf + 1 on [0, 2^70): squared spectral magnitudes incl. the tiny end
      const float x = __uint_as_float((h2 & 0x007fffffu) | (((h1 >> 7) % 140u + 57u) << 23));
      if (__float_as_uint(nsb_sqrtf_p1(x)) != __float_as_uint(__fadd_rn(__fsqrt_rn(x), 1.f))) ++mism;
      // nsb_sqrtf == sqrtf on [2^-49, 2^70) (energy ratios); 0 below 2^-50 by design
      if (x >= 1.7763568394002505e-15f && __float_as_uint(nsb_sqrtf(x)) != __float_as_uint(__fsqrt_rn(x))) ++mism;
      const float xs = __uint_as_float(h1 & 0x00ffffffu);   // subnormals and the smallest normals
      if (nsb_sqrtf(xs) != 0.f || nsb_sqrtf(0.f) != 0.f) ++mism;
      if (nsb_sqrtf_p1(xs) != 1.f || nsb_sqrtf_p1(0.f) != 1.f) ++mism;
    }
    {
      // packed fp32 (f32x2) forms of ns_warp.cuh against scalar arithmetic: a product that feeds a sum must not
      // be contracted into an FMA (ptxas does that to mul.f32x2 + add.f32x2 whatever --fmad says), the swapped /
      // negated second operands of FADD2 must mean what the source says, and a logarithm of either half of a
      // packed result must be that of the scalar (nvcc once dropped the exponent shift for one half)
      const float2 pa = make_float2(a, b), pb = make_float2(__uint_as_float((h2 & 0x807fffffu) | 0x3f000000u), a * 0.37f);
      const float2 pc = make_float2(b * 1.7f, __uint_as_float((h1 & 0x807fffffu) | 0x3f800000u));
      const float2 pd = make_float2(-b, 3.f * a);
      auto same2 = [](float2 g, float wx, float wy) { return __float_as_uint(g.x) == __float_as_uint(wx) && __float_as_uint(g.y) == __float_as_uint(wy); };
      if (!same2(vmadd(pa, pb, pc), __fadd_rn(__fmul_rn(pa.x, pb.x), pc.x), __fadd_rn(__fmul_rn(pa.y, pb.y), pc.y))) ++mism;
      if (!same2(vmmadd(pa, pb, pc, pd), __fadd_rn(__fmul_rn(pa.x, pb.x), __fmul_rn(pc.x, pd.x)),
                 __fadd_rn(__fmul_rn(pa.y, pb.y), __fmul_rn(pc.y, pd.y)))) ++mism;
      if (!same2(cmul(pb, pc), __fsub_rn(__fmul_rn(pb.x, pc.x), __fmul_rn(pb.y, pc.y)),
                 __fadd_rn(__fmul_rn(pb.x, pc.y), __fmul_rn(pb.y, pc.x)))) ++mism;
      if (!same2(cmul_conj(pb, pc), __fadd_rn(__fmul_rn(pb.x, pc.x), __fmul_rn(pb.y, pc.y)),
                 __fsub_rn(__fmul_rn(pb.y, pc.x), __fmul_rn(pb.x, pc.y)))) ++mism;
      if (!same2(cadd_i(pa, pc), __fsub_rn(pa.x, pc.y), __fadd_rn(pa.y, pc.x))) ++mism;
      if (!same2(csub_i(pa, pc), __fadd_rn(pa.x, pc.y), __fsub_rn(pa.y, pc.x))) ++mism;
      if (!same2(vfdiv(pa, make_float2(b, pc.y)), fdiv(pa.x, b), fdiv(pa.y, pc.y))) ++mism;
      const float2 sq = vmmadd(pb, pb, pc, pc);
      const float2 mg = vsqrt_p1(sq);
      if (!same2(mg, nsb_sqrtf_p1(__fadd_rn(__fmul_rn(pb.x, pb.x), __fmul_rn(pc.x, pc.x))),
                 nsb_sqrtf_p1(__fadd_rn(__fmul_rn(pb.y, pb.y), __fmul_rn(pc.y, pc.y))))) ++mism;
      const float2 lg = vlog_rn(vadd(mg, make_float2(fabsf(b), 0.25f)));
      if (!same2(lg, nsb_log_rn(__fadd_rn(mg.x, fabsf(b))), nsb_log_rn(__fadd_rn(mg.y, 0.25f)))) ++mism;
    }
    {
      // fx_udiv_q20 (ns_fixed.cuh): exact below its bound, never below the reference's cap above it; random
      // operands of every magnitude, and dividends one below / at / one above an exact multiple of the divisor
      const unsigned sh1 = (h1 >> 27), sh2 = (h2 >> 27);
      const unsigned bb = (h2 >> sh2) | 1u;
      unsigned aa = h1 >> sh1;
      for (int t = 0; t < 4; ++t) {
        if (t > 0) {
          const unsigned long long mlt = (unsigned long long)(h1 % 1200000u) * bb + (unsigned long long)(t - 2) ;
          if (mlt > 0xffffffffull) continue;
          aa = (unsigned)mlt;
          if (t == 1 && aa == 0xffffffffu) continue;
        }
        const unsigned want = aa / bb, got = fx_udiv_q20(aa, bb);
        if (want < kFxUdivExactBelow ? got != want : got < 1048575u) ++mism;
      }
    }
    const float c = (float)(h1 % 401u);   // small integers as in counters
    selftest_div(fdiv(c, (float)(h2 % 200u + 1u)), __fdiv_rn(c, (float)(h2 % 200u + 1u)), mism, ulp1, ndiv);
    int32_t v = (int32_t)h2, root = 0;
    for (int k = 15; k >= 0; --k) {
      const int32_t t = root + (1 << k);
      if (v >= (int32_t)((uint32_t)t << k)) { v -= (int32_t)((uint32_t)t << k); root |= 2 << k; }
    }
    if ((uint32_t)(root >> 1) != fx_sqrt_floor(h2)) ++mism;
  }
  if (mism) atomicAdd(out, mism);
  if (ulp1) atomicAdd(out + 1, ulp1);
  atomicAdd(out + 2, ndiv);
  if (lrn) atomicAdd(out + 3, lrn);
  atomicAdd(out + 4, nlog);
  if (ern) atomicAdd(out + 5, ern);
  atomicAdd(out + 6, nexp);
  if (trn) atomicAdd(out + 7, trn);
}

template <typename T>
int ProcessInterleaved(void* const* hv, int channels, T* data, int samples_per_channel);

__global__ void synth_kernel(int16_t* dst, size_t stride, int n_streams, uint32_t first_stream,
                             uint32_t fs, uint32_t first_sample, uint32_t n_samples, uint32_t seed) {
  // grid: x over sample pairs, y over streams (no per-element 64-bit index division)
  const uint32_t pairs = n_samples / 2;
  for (uint32_t s = blockIdx.y; s < (uint32_t)n_streams; s += gridDim.y)
  for (uint32_t pr = blockIdx.x * blockDim.x + threadIdx.x; pr < pairs; pr += gridDim.x * blockDim.x) {
    const uint32_t n = pr * 2u;
    // the generator divides 64-bit times by fs: with fs a literal those become multiplications
    uint32_t a, b;
#define NSB_SYNTH_PAIR(FS)                                                                   \
    a = (uint16_t)pcm_synth_sample(seed, first_stream + s, FS, first_sample + n);            \
    b = (uint16_t)pcm_synth_sample(seed, first_stream + s, FS, first_sample + n + 1u)
    if (fs == 16000u) { NSB_SYNTH_PAIR(16000u); }
    else if (fs == 8000u) { NSB_SYNTH_PAIR(8000u); }
    else if (fs == 32000u) { NSB_SYNTH_PAIR(32000u); }
    else if (fs == 48000u) { NSB_SYNTH_PAIR(48000u); }
    else { NSB_SYNTH_PAIR(fs); }
#undef NSB_SYNTH_PAIR
    reinterpret_cast<uint32_t*>(dst + (size_t)s * stride)[n / 2] = a | (b << 16);
  }
}

__global__ void checksum_kernel(const int16_t* pcm, size_t stride, int n_streams, uint32_t n_samples,
                                long long* sums, int accumulate) {
  const int s = blockIdx.x;
  if (s >= n_streams) return;
  long long a = 0, b = 0;
  const int16_t* p = pcm + (size_t)s * stride;
  for (uint32_t i = threadIdx.x; i < n_samples; i += blockDim.x) {
    const long long v = p[i];
    a += v;
    b += v * v;
  }
  __shared__ long long sa[256], sb[256];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums[2 * s] = (accumulate ? sums[2 * s] : 0) + sa[0];
    sums[2 * s + 1] = (accumulate ? sums[2 * s + 1] : 0) + sb[0];
  }
}

// One device: H2D interleaved -> deinterleave (+FloatToS16) -> split/NS/merge -> interleave
// (+S16ToFloat) -> D2H, in place (apm_ns.cpp:47-132 does the same per 10 ms on the host).
template <typename T>
int ProcessInterleaved(void* const* hv, int channels, T* data, int samples_per_channel) {
  UseLock lk;
  if (!data) return Fail("NULL data");
  std::vector<Handle*> hs;
  size_t no_stride_in = 0, no_stride_out = 0;
  if (CheckBatch(hv, channels, kMagicF, &no_stride_in, &no_stride_out, 0, &hs) != 0) return -1;
  const int fl = (int)hs[0]->fs / 100;
  if (samples_per_channel <= 0 || samples_per_channel % fl) return Fail("samples_per_channel must be a multiple of fs/100");
  if (t_memo.single_dev < 0) return Fail("channels of one capture stream must share a GPU");
  lk.Device(hs[0]->dev);
  DeviceCtx* d;
  if (DeviceReady(hs[0]->dev, &d) != 0) return -1;
  const int frames = samples_per_channel / fl;
  const size_t total = (size_t)channels * samples_per_channel;
  const size_t per = (size_t)samples_per_channel;   // fs/100 is a multiple of 8: rows stay 16-byte aligned
  T* d_il = nullptr;
  int16_t* d_pl = nullptr;
  CU_OK(cudaMalloc(&d_il, sizeof(T) * total));
  CU_OK(cudaMalloc(&d_pl, sizeof(int16_t) * per * channels));
  int rc = 0;
  do {
    cudaError_t e = cudaMemcpyAsync(d_il, data, sizeof(T) * total, cudaMemcpyHostToDevice, d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    deinterleave_kernel<T><<<grid, 256, 0, d->stream>>>(d_il, d_pl, channels, samples_per_channel);
    ++g_launches;
    if ((rc = RunDevice(*d, kMagicF, hs, d_pl, per, d_pl, per, frames, d->stream)) != 0) break;
    interleave_kernel<T><<<grid, 256, 0, d->stream>>>(d_pl, d_il, channels, samples_per_channel);
    ++g_launches;
    e = cudaMemcpyAsync(data, d_il, sizeof(T) * total, cudaMemcpyDeviceToHost, d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
    e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { rc = Fail(cudaGetErrorString(e)); break; }
  } while (0);
  cudaFree(d_il);
  cudaFree(d_pl);
  return rc;
}

}  // namespace
}  // namespace nsb200

using namespace nsb200;

// ---- stream state snapshot / restore / migration ----------------------
// A stream is its slabs (suppressor state, float histograms, band-split state) plus a few host
// fields; the blob is those verbatim behind a small header, so restore is bit-exact.
struct StateBlobHeader {
  uint32_t tag;          // 'NSB2'
  uint32_t magic;        // kMagicF / kMagicX
  uint32_t fs;
  int32_t mode;
  int32_t init_flag;
  int32_t analyze_seen;
  int32_t split_mode;
  int32_t pad0;
  double down_vsi;
  uint32_t state_bytes, hist_bytes, band_bytes;
  uint32_t reserved;
  float analyze_frame[160];
};
constexpr uint32_t kBlobTag = 0x3242534eu;

Handle* AnyHandle(const void* hv) {
  Handle* h = static_cast<Handle*>(const_cast<void*>(hv));
  return (h && (h->magic == kMagicF || h->magic == kMagicX)) ? h : nullptr;
}
struct SlabRefs { void* ptr[3]; size_t bytes[3]; };
SlabRefs SlabsOf(DeviceCtx& d, const Handle* h, int slot) {
  SlabRefs r = {};
  SlabPool& sp = h->magic == kMagicF ? d.f_state : d.x_state;
  r.ptr[0] = static_cast<char*>(sp.base) + (size_t)slot * sp.slab_bytes;
  r.bytes[0] = sp.slab_bytes;
  if (h->magic == kMagicF) {
    r.ptr[1] = static_cast<char*>(d.f_hist.base) + (size_t)slot * d.f_hist.slab_bytes;
    r.bytes[1] = d.f_hist.slab_bytes;
  }
  const int bslot = h->magic == kMagicF ? 2 * slot : 2 * slot + 1;
  r.ptr[2] = static_cast<char*>(d.b_state.base) + (size_t)bslot * d.b_state.slab_bytes;
  r.bytes[2] = d.b_state.slab_bytes;
  return r;
}
size_t StateSize(const void* hv) {
  UseLock lk;
  Handle* h = AnyHandle(hv);
  if (!h) return 0;
  SlabRefs r = SlabsOf(g_devs[h->dev], h, h->slot);
  return sizeof(StateBlobHeader) + r.bytes[0] + r.bytes[1] + r.bytes[2];
}
int ExportState(const void* hv, void* buf, size_t size) {
  UseLock lk;
  Handle* h = AnyHandle(hv);
  if (!h) return Fail("bad handle");
  if (!buf) return Fail("NULL buffer");
  lk.Device(h->dev);
  DeviceCtx* d;
  if (DeviceReady(h->dev, &d) != 0) return -1;
  SlabRefs r = SlabsOf(*d, h, h->slot);
  const size_t need = sizeof(StateBlobHeader) + r.bytes[0] + r.bytes[1] + r.bytes[2];
  if (size < need) return Fail("state buffer too small");
  CU_OK(cudaDeviceSynchronize());   // everything enqueued for this stream has landed
  StateBlobHeader hd = {};
  hd.tag = kBlobTag;
  hd.magic = h->magic;
  hd.fs = h->fs;
  hd.mode = h->mode;
  hd.init_flag = h->init_flag;
  hd.analyze_seen = h->analyze_seen ? 1 : 0;
  hd.split_mode = h->split_mode ? 1 : 0;
  hd.down_vsi = h->down_vsi;
  hd.state_bytes = (uint32_t)r.bytes[0];
  hd.hist_bytes = (uint32_t)r.bytes[1];
  hd.band_bytes = (uint32_t)r.bytes[2];
  memcpy(hd.analyze_frame, h->analyze_frame, sizeof(hd.analyze_frame));
  char* p = static_cast<char*>(buf);
  memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  for (int k = 0; k < 3; ++k) {
    if (r.bytes[k]) CU_OK(cudaMemcpy(p, r.ptr[k], r.bytes[k], cudaMemcpyDeviceToHost));
    p += r.bytes[k];
  }
  return 0;
}
// What a blob may not carry: the kernels trust the slab's control words.  (The float kernel indexes
// nothing with them -- its histogram bins are range-checked feature values -- but its counters steer
// divisions; the fixed-point kernel looks up 1 / (counter + 1) in a 201-entry table.)
bool BlobHeaderSane(const StateBlobHeader& hd, const char* slab) {
  if (!(hd.fs == 8000 || hd.fs == 16000 || hd.fs == 32000 || hd.fs == 48000)) return false;
  if (hd.mode < 0 || hd.mode > 3 || (hd.init_flag != 0 && hd.init_flag != 1)) return false;
  if (!(hd.down_vsi >= 0.0 && hd.down_vsi < 4096.0)) return false;
  if (!hd.init_flag) return true;
  int32_t w[kNsfHdrWords > kNsxHdrWords ? kNsfHdrWords : kNsxHdrWords];
  memcpy(w, slab, sizeof(w));
  if (hd.magic == kMagicF) {
    if (w[kH_blockInd] < -1 || w[kH_updates] < 0 || w[kH_updates] > 200) return false;
    for (int s = 0; s < 3; ++s)
      if (w[kH_counter + s] < 0 || w[kH_counter + s] > 200) return false;
    if (w[kH_modelUpd0] < 0 || w[kH_modelUpd0] > 2 || w[kH_modelUpd3] < 0 || w[kH_modelUpd3] > 500) return false;
    if ((uint32_t)w[kH_fs] != hd.fs || w[kH_mode] != hd.mode) return false;
    return true;
  }
  return nsx_header_sane(reinterpret_cast<const uint32_t*>(w), hd.fs, hd.mode);
}
int ImportState(void* hv, const void* buf, size_t size) {
  ExclusiveLock lk;
  ++g_epoch;   // handle lists validated so far are stale (BatchMemo)
  Handle* h = AnyHandle(hv);
  if (!h) return Fail("bad handle");
  if (!buf || size < sizeof(StateBlobHeader)) return Fail("state blob too small");
  StateBlobHeader hd;
  memcpy(&hd, buf, sizeof(hd));
  if (hd.tag != kBlobTag) return Fail("not a state blob");
  if (hd.magic != h->magic) return Fail("state blob is of the other suppressor kind (float / fixed)");
  DeviceCtx* d;
  if (DeviceReady(h->dev, &d) != 0) return -1;
  SlabRefs r = SlabsOf(*d, h, h->slot);
  if (hd.state_bytes != r.bytes[0] || hd.hist_bytes != r.bytes[1] || hd.band_bytes != r.bytes[2])
    return Fail("state blob from an incompatible library version");
  if (size < sizeof(hd) + r.bytes[0] + r.bytes[1] + r.bytes[2]) return Fail("state blob truncated");
  if (!BlobHeaderSane(hd, static_cast<const char*>(buf) + sizeof(hd))) return Fail("state blob carries control words out of range");
  CU_OK(cudaDeviceSynchronize());
  const char* p = static_cast<const char*>(buf) + sizeof(hd);
  for (int k = 0; k < 3; ++k) {
    if (r.bytes[k]) CU_OK(cudaMemcpy(r.ptr[k], p, r.bytes[k], cudaMemcpyHostToDevice));
    p += r.bytes[k];
  }
  h->fs = hd.fs;
  h->mode = hd.mode;
  h->init_flag = hd.init_flag;
  h->analyze_seen = hd.analyze_seen != 0;
  h->split_mode = hd.split_mode != 0;
  h->down_vsi = hd.down_vsi;
  memcpy(h->analyze_frame, hd.analyze_frame, sizeof(h->analyze_frame));
  return 0;
}
// Moves a stream to another GPU: new slot there, slabs copied device to device (NVLink peer copy
// when the GPUs are peers, staged by the driver otherwise), old slot released.
int MigrateHandle(void* hv, int device) {
  ExclusiveLock lk;
  ++g_epoch;   // handle lists validated so far are stale (BatchMemo)
  Handle* h = AnyHandle(hv);
  if (!h) return Fail("bad handle");
  if (EnsureDevices() != 0) return -1;
  if (device < 0 || device >= (int)g_devs.size()) return Fail("bad device index");
  if (device == h->dev) return 0;
  DeviceCtx *src, *dst;
  if (DeviceReady(h->dev, &src) != 0) return -1;
  CU_OK(cudaDeviceSynchronize());
  if (DeviceReady(device, &dst) != 0) return -1;
  CU_OK(cudaDeviceSynchronize());
  int slot = -1;
  int rc;
  if (h->magic == kMagicF) {
    rc = PoolAlloc(*dst, dst->f_state, &slot);
    if (rc == 0 && PoolGrow(*dst, dst->f_hist, dst->f_state.capacity) != 0) rc = -1;
  } else {
    rc = PoolAlloc(*dst, dst->x_state, &slot);
  }
  if (rc == 0 && PoolGrow(*dst, dst->b_state, 2 * (h->magic == kMagicF ? dst->f_state : dst->x_state).capacity) != 0) rc = -1;
  if (rc != 0) return -1;
  SlabRefs a = SlabsOf(*src, h, h->slot), b = SlabsOf(*dst, h, slot);
  for (int k = 0; k < 3; ++k)
    if (a.bytes[k]) CU_OK(cudaMemcpyPeer(b.ptr[k], device, a.ptr[k], h->dev, a.bytes[k]));
  (h->magic == kMagicF ? src->f_state : src->x_state).free_slots.push_back(h->slot);
  src->cached_slots.clear();
  dst->cached_slots.clear();
  h->dev = device;
  h->slot = slot;
  return 0;
}

extern "C" {

int WebRtcNs_Create(NsHandle** h) { return Create(reinterpret_cast<void**>(h), kMagicF); }
int WebRtcNs_Free(NsHandle* h) { return Free(h, kMagicF); }
int WebRtcNs_Init(NsHandle* h, uint32_t fs) {
  void* hv = h;
  return InitMany(&hv, 1, fs, 0, kMagicF);
}
int WebRtcNs_set_policy(NsHandle* h, int mode) { return SetPolicy(h, mode, kMagicF); }

// WebRtcNs_Analyze updates the statistics itself, as in the reference (ns_core.c:1043-1181): the Analyze half
// of the frame runs on the GPU at once, so that WebRtcNs_prior_speech_probability read between Analyze and
// Process (noise_suppression.c:57-66) is current; WebRtcNs_Process then runs the Process half on its own frame
// (which may differ from the analysed one: audio_processing_impl.cc:625-631).  Process without a preceding
// Analyze feeds its frame to both halves, as every caller in the reference tree does anyway.
void WebRtcNs_Analyze(NsHandle* hv, const float* spframe) {
  Handle* h = AsHandle(hv, kMagicF);
  if (!h || !h->init_flag || !spframe) {
    Fail("WebRtcNs_Analyze: handle not initialised");
    return;
  }
  const int fl = h->fs == 8000 ? 80 : 160;
  float frame[160];
  memcpy(frame, spframe, sizeof(float) * fl);
  void* one = h;
  if (ProcessBandsF32(&one, 1, 1, nullptr, 0, nullptr, 0, 1, frame, (size_t)fl, 1) == 0) h->analyze_seen = true;
}

void WebRtcNs_Process(NsHandle* hv, const float* const* spframe, int num_bands, float* const* outframe) {
  Handle* h = AsHandle(hv, kMagicF);
  if (!h || !h->init_flag || !spframe || !outframe || num_bands < 1 || num_bands > 3) {
    Fail("WebRtcNs_Process: handle not initialised or bad arguments");
    return;
  }
  const int fl = h->fs == 8000 ? 80 : 160;
  float in[3 * 160], out[3 * 160];
  for (int b = 0; b < num_bands; ++b) memcpy(in + b * fl, spframe[b], sizeof(float) * fl);
  const bool analysed = h->analyze_seen;
  h->analyze_seen = false;
  void* one = h;
  if (ProcessBandsF32(&one, 1, num_bands, in, (size_t)num_bands * fl, out, (size_t)num_bands * fl, 1, nullptr, 0,
                      analysed ? 2 : 0) != 0)
    return;
  for (int b = 0; b < num_bands; ++b) memcpy(outframe[b], out + b * fl, sizeof(float) * fl);
}

float WebRtcNs_prior_speech_probability(NsHandle* hv) {
  UseLock lk;
  Handle* h = AsHandle(hv, kMagicF);
  if (!h || !h->init_flag) return -1.f;
  lk.Device(h->dev);
  DeviceCtx* d;
  if (DeviceReady(h->dev, &d) != 0) return -1.f;
  float v = -1.f;
  cudaStreamSynchronize(d->stream);
  if (cudaMemcpy(&v, (float*)d->f_state.base + (size_t)h->slot * kNsfStateWords + kH_priorSpeechProb,
                 sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1.f;
  return v;
}

int WebRtcNsx_Create(NsxHandle** h) { return Create(reinterpret_cast<void**>(h), kMagicX); }
int WebRtcNsx_Free(NsxHandle* h) { return Free(h, kMagicX); }
int WebRtcNsx_Init(NsxHandle* h, uint32_t fs) {
  void* hv = h;
  return InitMany(&hv, 1, fs, 0, kMagicX);
}
int WebRtcNsx_set_policy(NsxHandle* h, int mode) { return SetPolicy(h, mode, kMagicX); }

void WebRtcNsx_Process(NsxHandle* hv, const short* const* speechFrame, int num_bands, short* const* outFrame) {
  Handle* h = AsHandle(hv, kMagicX);
  if (!h || !h->init_flag || !speechFrame || !outFrame || num_bands < 1 || num_bands > 3) {
    Fail("WebRtcNsx_Process: handle not initialised or bad arguments");
    return;
  }
  UseLock lk;
  lk.Device(h->dev);
  const int fl = h->fs == 8000 ? 80 : 160;
  const int ana = h->fs == 8000 ? 128 : 256;
  DeviceCtx* d;
  if (DeviceReady(h->dev, &d) != 0) return;
  int16_t buf[3 * 160];
  for (int b = 0; b < num_bands; ++b) memcpy(buf + b * fl, speechFrame[b], sizeof(int16_t) * fl);
  int16_t* dio = nullptr;
  if (cudaMalloc(&dio, sizeof(buf)) != cudaSuccess) { Fail("cudaMalloc"); return; }
  std::vector<int> slots(1, h->slot);
  if (UploadSlots(*d, slots, d->stream) == 0 &&
      cudaMemcpyAsync(dio, buf, sizeof(int16_t) * num_bands * fl, cudaMemcpyHostToDevice, d->stream) == cudaSuccess) {
    NsxLaunch p;
    p.state = (uint32_t*)d->x_state.base;
    p.slots = d->d_slots;
    p.tables = d->d_nsx_tables;
    p.in = dio;
    p.out = dio;
    p.in_stream_stride = p.out_stream_stride = num_bands * fl;
    p.in_frame_stride = p.out_frame_stride = num_bands * fl;
    p.in_band_stride = p.out_band_stride = fl;
    p.n_streams = 1;
    p.frames = 1;
    if (LaunchNsx(ana, num_bands, p, d->stream) == 0 &&
        cudaMemcpyAsync(buf, dio, sizeof(int16_t) * num_bands * fl, cudaMemcpyDeviceToHost, d->stream) == cudaSuccess &&
        cudaStreamSynchronize(d->stream) == cudaSuccess) {
      for (int b = 0; b < num_bands; ++b) memcpy(outFrame[b], buf + b * fl, sizeof(int16_t) * fl);
    } else {
      Fail("WebRtcNsx_Process: launch or copy failed");
    }
  }
  cudaFree(dio);
}

int WebRtcNs_ProcessBatch(NsHandle* const* hs, int n, const int16_t* in, size_t is, int16_t* out, size_t os, int frames) {
  return BatchHost(reinterpret_cast<void* const*>(hs), n, kMagicF, in, is, out, os, frames);
}
int WebRtcNsx_ProcessBatch(NsxHandle* const* hs, int n, const int16_t* in, size_t is, int16_t* out, size_t os, int frames) {
  return BatchHost(reinterpret_cast<void* const*>(hs), n, kMagicX, in, is, out, os, frames);
}
int WebRtcNs_ProcessBatchAsync(NsHandle* const* hs, int n, const int16_t* in, size_t is, int16_t* out, size_t os, int frames,
                               uint64_t* ticket) {
  if (!ticket) return Fail("NULL ticket pointer");
  return BatchHost(reinterpret_cast<void* const*>(hs), n, kMagicF, in, is, out, os, frames, ticket);
}
int WebRtcNsx_ProcessBatchAsync(NsxHandle* const* hs, int n, const int16_t* in, size_t is, int16_t* out, size_t os, int frames,
                                uint64_t* ticket) {
  if (!ticket) return Fail("NULL ticket pointer");
  return BatchHost(reinterpret_cast<void* const*>(hs), n, kMagicX, in, is, out, os, frames, ticket);
}
int WebRtcNsB200_WaitBatch(uint64_t ticket) { return WaitBatch(ticket); }
int WebRtcNs_ProcessBatchDevice(NsHandle* const* hs, int n, const int16_t* in, size_t is, int16_t* out, size_t os,
                                int frames, void* st) {
  return BatchDevice(reinterpret_cast<void* const*>(hs), n, kMagicF, in, is, out, os, frames, st);
}
int WebRtcNsx_ProcessBatchDevice(NsxHandle* const* hs, int n, const int16_t* in, size_t is, int16_t* out, size_t os,
                                 int frames, void* st) {
  return BatchDevice(reinterpret_cast<void* const*>(hs), n, kMagicX, in, is, out, os, frames, st);
}
int WebRtcNs_ProcessBatchBandsF32(NsHandle* const* hs, int n, int nb, const float* in, size_t is, float* out,
                                  size_t os, int frames) {
  return ProcessBandsF32(reinterpret_cast<void* const*>(hs), n, nb, in, is, out, os, frames);
}
int WebRtcNs_AnalyzeProcessBatch(NsHandle* const* hs, int n, const int16_t* ana, size_t as, const int16_t* in, size_t is,
                                 int16_t* out, size_t os, int frames) {
  return SplitBatchHost((void* const*)hs, n, ana, as, in, is, out, os, frames);
}
int WebRtcNs_AnalyzeProcessBatchDevice(NsHandle* const* hs, int n, const int16_t* ana, size_t as, const int16_t* in,
                                       size_t is, int16_t* out, size_t os, int frames, void* stream) {
  return SplitBatchDevice((void* const*)hs, n, ana, as, in, is, out, os, frames, stream);
}
int WebRtcNs_AnalyzeProcessBatchBandsF32(NsHandle* const* hs, int n, int nb, const float* ana, size_t as, const float* in,
                                         size_t is, float* out, size_t os, int frames) {
  if (!ana) return Fail("NULL Analyze signal");
  return ProcessBandsF32((void* const*)hs, n, nb, in, is, out, os, frames, ana, as);
}
int WebRtcNs_ProcessInterleavedI16(NsHandle* const* hs, int n_channels, int16_t* data, int samples_per_channel) {
  return ProcessInterleaved<int16_t>(reinterpret_cast<void* const*>(hs), n_channels, data, samples_per_channel);
}
int WebRtcNs_ProcessInterleavedF32(NsHandle* const* hs, int n_channels, float* data, int samples_per_channel) {
  return ProcessInterleaved<float>(reinterpret_cast<void* const*>(hs), n_channels, data, samples_per_channel);
}
int WebRtcNs_InitBatch(NsHandle* const* hs, int n, uint32_t fs, int mode) {
  return InitMany(reinterpret_cast<void* const*>(hs), n, fs, mode, kMagicF);
}
int WebRtcNsx_InitBatch(NsxHandle* const* hs, int n, uint32_t fs, int mode) {
  return InitMany(reinterpret_cast<void* const*>(hs), n, fs, mode, kMagicX);
}

size_t WebRtcNsB200_StateSize(const void* handle) { return StateSize(handle); }
int WebRtcNsB200_ExportState(const void* handle, void* buf, size_t size) { return ExportState(handle, buf, size); }
int WebRtcNsB200_ImportState(void* handle, const void* buf, size_t size) { return ImportState(handle, buf, size); }
int WebRtcNsB200_MigrateHandle(void* handle, int device) { return MigrateHandle(handle, device); }
int WebRtcNsB200_HandleDevice(const void* handle) {
  UseLock lk;
  Handle* h = AnyHandle(handle);
  return h ? h->dev : -1;
}

int WebRtcNsB200_SetCreateDevice(int device) {   // per calling thread
  if (device >= 0) {
    if (EnsureDevices() != 0) return -1;
    if (device >= (int)g_devs.size()) return Fail("bad device index");
  }
  t_create_device = device;
  return 0;
}
int WebRtcNsB200_DeviceCount(void) {
  if (EnsureDevices() != 0) return 0;
  return (int)g_devs.size();
}
int WebRtcNsB200_Synchronize(void) {
  ExclusiveLock lk;
  for (auto& d : g_devs) {
    if (!d.ready) continue;
    CU_OK(cudaSetDevice(d.dev));
    CU_OK(cudaDeviceSynchronize());
  }
  return 0;
}
const char* WebRtcNsB200_LastError(void) { return t_err.c_str(); }   // of the calling thread
uint64_t WebRtcNsB200_KernelLaunches(void) { return g_launches; }

int WebRtcNsB200_SelfTestStats(uint64_t n_cases, uint64_t* stats) {
  UseLock lk;
  int dev = 0;
  if (EnsureDevices() != 0) return -1;
  cudaGetDevice(&dev);
  lk.Device(dev);
  DeviceCtx* d;
  if (DeviceReady(dev, &d) != 0) return -1;
  unsigned long long* bad = nullptr;
  CU_OK(cudaMalloc(&bad, 8 * sizeof(*bad)));
  CU_OK(cudaMemsetAsync(bad, 0, 8 * sizeof(*bad), d->stream));
  selftest_kernel<<<148 * 4, 256, 0, d->stream>>>(n_cases, 12345u, bad);
  ++g_launches;
  unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CU_OK(cudaMemcpyAsync(h, bad, sizeof(h), cudaMemcpyDeviceToHost, d->stream));
  CU_OK(cudaStreamSynchronize(d->stream));
  cudaFree(bad);
  for (int i = 0; i < 8; ++i) stats[i] = h[i];
  return 0;
}
int WebRtcNsB200_SelfTest(uint64_t n_cases) {
  uint64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (WebRtcNsB200_SelfTestStats(n_cases, st) != 0) return -1;
  if (st[0] != 0) return Fail("self-test: " + std::to_string(st[0]) + " arithmetic mismatches");
  if (st[3] * 100000ull > st[4])
    return Fail("self-test: " + std::to_string(st[3]) + " of " + std::to_string(st[4]) + " logarithms not the rounded double-precision one");
  if (st[5] * 100000ull > st[6] || st[7] * 100000ull > st[6])
    return Fail("self-test: " + std::to_string(st[5]) + " exponentials / " + std::to_string(st[7]) + " sigmoid maps of " +
                std::to_string(st[6]) + " not the rounded double-precision ones");
  // fdiv(): correctly rounded except for a handful of near-halfway quotients (ns_warp.cuh)
  if (st[1] * 1000000ull > st[2] * 2ull)
    return Fail("self-test: " + std::to_string(st[1]) + " of " + std::to_string(st[2]) + " divisions one ulp off");
  return 0;
}

int WebRtcNsB200_SynthPcmDevice(int16_t* dst, size_t stride, int n_streams, uint32_t first_stream, uint32_t fs,
                                uint32_t first_sample, uint32_t n_samples, uint32_t seed, void* st) {
  if ((stride & 1) || (n_samples & 1)) return Fail("stride and n_samples must be even");
  if (n_streams <= 0 || n_samples == 0) return 0;
  unsigned gx = (n_samples / 2 + 255) / 256;
  if (gx > 32) gx = 32;
  const dim3 grid(gx, n_streams < 65535 ? n_streams : 65535);
  synth_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(dst, stride, n_streams, first_stream, fs, first_sample,
                                                   n_samples, seed);
  ++g_launches;
  CU_OK(cudaGetLastError());
  return 0;
}
void WebRtcNsB200_SynthPcmHost(int16_t* dst, uint32_t stream, uint32_t fs, uint32_t first_sample,
                               uint32_t n_samples, uint32_t seed) {
  for (uint32_t i = 0; i < n_samples; ++i) dst[i] = pcm_synth_sample(seed, stream, fs, first_sample + i);
}
int WebRtcNsB200_ChecksumDevice(const int16_t* pcm, size_t stride, int n_streams, uint32_t n_samples,
                                int64_t* sums, void* st) {
  checksum_kernel<<<n_streams, 256, 0, (cudaStream_t)st>>>(pcm, stride, n_streams, n_samples, (long long*)sums, 0);
  ++g_launches;
  CU_OK(cudaGetLastError());
  return 0;
}
int WebRtcNsB200_ChecksumAccumulateDevice(const int16_t* pcm, size_t stride, int n_streams, uint32_t n_samples,
                                          int64_t* sums, void* st) {
  checksum_kernel<<<n_streams, 256, 0, (cudaStream_t)st>>>(pcm, stride, n_streams, n_samples, (long long*)sums, 1);
  ++g_launches;
  CU_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
