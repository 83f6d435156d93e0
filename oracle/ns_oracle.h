/* TEST INFRASTRUCTURE ONLY.  Public face of liboracle_ns.so: a scalar CPU
 * restatement of the reference's noise-suppression path, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker --
 * never by the product. */
#ifndef ORACLE_NS_ORACLE_H_
#define ORACLE_NS_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- float suppressor (WebRtcNs_*) ---- */
typedef struct NsfOracle NsfOracle;
NsfOracle* nsf_oracle_create(void);
void nsf_oracle_free(NsfOracle* s);
int nsf_oracle_init(NsfOracle* s, uint32_t fs);                 /* 0 / -1 */
int nsf_oracle_set_policy(NsfOracle* s, int mode);              /* 0 / -1 */
void nsf_oracle_analyze(NsfOracle* s, const float* frame);
void nsf_oracle_process(NsfOracle* s, const float* const* in, int num_bands, float* const* out);
float nsf_oracle_prior_speech_probability(const NsfOracle* s);  /* -1 when uninitialised */
/* one stream, 8/16 kHz: Analyze + Process per frame; out_f32 in int16 scale */
int nsf_oracle_run(int fs, int mode, int nframes, const int16_t* pcm_in, float* out_f32, float* prior_prob);

/* ---- fixed-point suppressor (WebRtcNsx_*) ---- */
typedef struct NsxOracle NsxOracle;
NsxOracle* nsx_oracle_create(void);
void nsx_oracle_free(NsxOracle* s);
int nsx_oracle_init(NsxOracle* s, uint32_t fs);                 /* 0 / -1 */
int nsx_oracle_set_policy(NsxOracle* s, int mode);              /* 0 / -1 */
void nsx_oracle_process(NsxOracle* s, const int16_t* const* in, int num_bands, int16_t* const* out);
/* one stream, 8/16 kHz, nframes frames of fs/100 samples */
int nsx_oracle_run(int fs, int mode, int nframes, const int16_t* pcm_in, int16_t* pcm_out);
/* primitives for known-answer tests */
int nsx_oracle_real_fft(int order, int inverse, const int16_t* in, int16_t* out);
const int16_t* nsx_oracle_table(const char* name, int* len);

/* the float path's real FFT alone (n = 256 / 128; isgn >= 0 forward, < 0 backward unscaled), in place */
int nsf_oracle_rdft(int n, int isgn, float* a);

/* ---- 32/48 kHz band split / merge (AudioBuffer + SplittingFilter) ---- */
typedef struct BandOracle BandOracle;
BandOracle* band_oracle_create(int fs);
void band_oracle_free(BandOracle* b);
int band_oracle_num_bands(const BandOracle* b);
void band_oracle_split(BandOracle* b, const int16_t* in, int16_t* bands);   /* bands [nb][160] */
void band_oracle_merge(BandOracle* b, const int16_t* bands, int16_t* out);
void band_oracle_qmf_analysis(const int16_t* in, int len, int16_t* low, int16_t* high, int32_t* st1, int32_t* st2);
void band_oracle_qmf_synthesis(const int16_t* low, const int16_t* high, int band_len, int16_t* out,
                               int32_t* st1, int32_t* st2);
/* whole path for one stream at any supported rate: split -> NSx (fixed=1) / float NS -> merge */
int band_oracle_run(int fixed, int fs, int mode, int nframes, const int16_t* pcm_in, int16_t* pcm_out);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_NS_ORACLE_H_ */
