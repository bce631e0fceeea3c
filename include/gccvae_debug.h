/* Development aids of libgccvae.so: pipeline timelines, stream markers, a TMA box probe.  NOT part of the drop-in
 * boundary (include/gccvae.h): nothing on the product path calls them; scripts/timeline_probe.py, scripts/graph_timeline.py
 * and the TMA layout test (tests/test_gpu_tc.py) do. */
#ifndef GCCVAE_DEBUG_H
#define GCCVAE_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif

/* debug aid: while set (device buffer of 32*8 int64), block 0 of every tap-GEMM launch records clock64()
 * at its pipeline events: [item][0 slot free,1 TMA issued,2 TMEM free,3 operands landed,4 accum ready,
 * 5 accum read,6 stored]. */
void gccvae_debug_set_timeline(long long* dev_buf);
/* debug: a 1-thread kernel that stores %globaltimer (ns) into buf[idx] when the stream reaches it */
int gccvae_debug_mark(long long* buf, int idx, void* stream);

/* debug aid: one 4-D TMA box load of a bf16 NHWC tensor, raw shared-memory image copied to `out`. */
int gccvae_debug_tma4d(const void* src_bf16, int N, int H, int W, int C, int kc, int bw, int bh, int bn, int es,
                       int c0, int c1, int c2, int c3, void* out, int out_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
