// Linux clone of the reference's command-line driver WebRtc_AMP_Port/test_ns_module.cpp:
//   test_ns_module in.wav out.wav [--fixed] [--mode N] [--block FRAMES]
// reads a 16-bit PCM WAV (mono or interleaved multi-channel, 8/16/32/48 kHz), runs the noise
// suppressor with policy kModerate (1) like the reference driver (test_ns_module.cpp:71) and
// writes the result with the same header.  Float NS by default (Analyze + Process per 10 ms,
// band split at 32/48 kHz: test_ns_module.cpp:94-103); --fixed = the -DNS_FIXED build
// (WebRtcNsx_Process, :90-92).  All processing happens in libwebrtc_ns_b200.so on the GPU;
// --block N hands N frames per call to the batched entry point (default 1 = the reference's loop).
// WAV parsing follows WebRtc_AMP_Port/wav_io.c:32-128 (RIFF/fmt/data chunks, 16-bit only).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/webrtc_ns_b200.h"

struct WavInfo {
  uint16_t channels, bits;
  uint32_t rate, data_bytes;
  long data_pos;
  std::vector<uint8_t> header;   // everything up to the sample data, copied to the output
};

static bool ReadHeader(FILE* f, WavInfo* w) {
  uint8_t b[12];
  if (fread(b, 1, 12, f) != 12 || memcmp(b, "RIFF", 4) || memcmp(b + 8, "WAVE", 4)) return false;
  bool have_fmt = false;
  for (;;) {
    uint8_t ch[8];
    if (fread(ch, 1, 8, f) != 8) return false;
    uint32_t sz = ch[4] | (ch[5] << 8) | (ch[6] << 16) | ((uint32_t)ch[7] << 24);
    if (!memcmp(ch, "fmt ", 4)) {
      std::vector<uint8_t> fmt(sz);
      if (fread(fmt.data(), 1, sz, f) != sz || sz < 16) return false;
      if ((fmt[0] | (fmt[1] << 8)) != 1) return false;           // PCM only
      w->channels = fmt[2] | (fmt[3] << 8);
      w->rate = fmt[4] | (fmt[5] << 8) | (fmt[6] << 16) | ((uint32_t)fmt[7] << 24);
      w->bits = fmt[14] | (fmt[15] << 8);
      have_fmt = true;
    } else if (!memcmp(ch, "data", 4)) {
      w->data_bytes = sz;
      w->data_pos = ftell(f);
      break;
    } else {
      fseek(f, (long)sz + (sz & 1), SEEK_CUR);
    }
  }
  if (!have_fmt) return false;
  w->header.resize((size_t)w->data_pos);
  fseek(f, 0, SEEK_SET);
  if (fread(w->header.data(), 1, w->header.size(), f) != w->header.size()) return false;
  return true;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s in.wav out.wav [--fixed] [--mode N] [--block FRAMES]\n", argv[0]);
    return -1;
  }
  bool fixed = false;
  int mode = 1, block = 1;
  for (int i = 3; i < argc; ++i) {
    if (!strcmp(argv[i], "--fixed")) fixed = true;
    else if (!strcmp(argv[i], "--mode") && i + 1 < argc) mode = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--block") && i + 1 < argc) block = atoi(argv[++i]);
  }
  FILE* fr = fopen(argv[1], "rb");
  FILE* fw = fopen(argv[2], "wb");
  if (!fr || !fw) {
    fprintf(stderr, "Fail to open file!!!\n");
    return -1;
  }
  WavInfo w;
  if (!ReadHeader(fr, &w)) {
    fprintf(stderr, "Fail to read wav header!\n");
    return -1;
  }
  if (w.bits != 16) {
    fprintf(stderr, "Now only support 16 bits per sample!\n");
    return -1;
  }
  printf("Process %s -> %s: %u Hz, %u channel(s), %s NS, policy %d\n", argv[1], argv[2], w.rate, w.channels,
         fixed ? "fixed" : "float", mode);
  const int ch = w.channels, fl = (int)w.rate / 100;
  std::vector<void*> hs(ch);
  for (int c = 0; c < ch; ++c) {
    int rc = fixed ? WebRtcNsx_Create((NsxHandle**)&hs[c]) : WebRtcNs_Create((NsHandle**)&hs[c]);
    if (rc == 0) rc = fixed ? WebRtcNsx_Init((NsxHandle*)hs[c], w.rate) : WebRtcNs_Init((NsHandle*)hs[c], w.rate);
    if (rc == 0) rc = fixed ? WebRtcNsx_set_policy((NsxHandle*)hs[c], mode) : WebRtcNs_set_policy((NsHandle*)hs[c], mode);
    if (rc != 0) {
      fprintf(stderr, "NS init failed: %s\n", WebRtcNsB200_LastError());
      return -1;
    }
  }
  fwrite(w.header.data(), 1, w.header.size(), fw);
  fseek(fr, w.data_pos, SEEK_SET);
  const size_t frame_samples = (size_t)fl * ch;
  std::vector<int16_t> buf(frame_samples * block), planar(frame_samples * block);
  size_t total_frames = w.data_bytes / (2 * frame_samples), done = 0;
  while (done < total_frames) {
    const int nf = (int)((total_frames - done) < (size_t)block ? (total_frames - done) : (size_t)block);
    if (fread(buf.data(), 2, frame_samples * nf, fr) != frame_samples * nf) break;
    int rc;
    if (!fixed) {
      rc = WebRtcNs_ProcessInterleavedI16((NsHandle* const*)hs.data(), ch, buf.data(), nf * fl);
    } else {
      // the -DNS_FIXED driver hands the undivided frame to WebRtcNsx_Process (test_ns_module.cpp:90-92);
      // here the batched call also performs the band split a correct 32/48 kHz caller needs
      for (int c = 0; c < ch; ++c)
        for (int n = 0; n < nf * fl; ++n) planar[(size_t)c * nf * fl + n] = buf[(size_t)n * ch + c];
      rc = WebRtcNsx_ProcessBatch((NsxHandle* const*)hs.data(), ch, planar.data(), (size_t)nf * fl, planar.data(),
                                  (size_t)nf * fl, nf);
      for (int c = 0; c < ch; ++c)
        for (int n = 0; n < nf * fl; ++n) buf[(size_t)n * ch + c] = planar[(size_t)c * nf * fl + n];
    }
    if (rc != 0) {
      fprintf(stderr, "processing failed: %s\n", WebRtcNsB200_LastError());
      return -1;
    }
    fwrite(buf.data(), 2, frame_samples * nf, fw);
    done += nf;
  }
  // trailing partial frame (and any chunks after "data") are copied through unchanged
  int c;
  while ((c = fgetc(fr)) != EOF) fputc(c, fw);
  printf("%zu frames\n", done);
  for (int i = 0; i < ch; ++i) fixed ? WebRtcNsx_Free((NsxHandle*)hs[i]) : WebRtcNs_Free((NsHandle*)hs[i]);
  fclose(fr);
  fclose(fw);
  return 0;
}
