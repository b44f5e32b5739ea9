// Counter-based dropout masks (Philox4x32-10) shared by every kernel that applies or re-applies one.
//
// The reference trains with nn.Dropout at three sites (hybrid_model.py:47,58,67-73,108; model.py:27,33-42):
// after GCN layers 1-3 (all four in STGCN.forward), between the LSTM layers, and on the head input.  torch's
// generator stream cannot be matched by a kernel that batches nodes, windows and tasks, so the
// contract here is the distribution, not the stream: element e of site s in forward pass c is kept with
// probability 1 - p and scaled by 1 / (1 - p), decided by one 32-bit word of
//     philox4x32_10(counter = {e / 4 (64 bit), s, c}, key = seed)[e % 4].
// Nothing is stored: the backward kernels regenerate the same words from (seed, pass, site, element).
// e is the element's index in the site's canonical row-major tensor:
//     GCN layer i   (site i)        e = (window-global row) * C + channel           over [G*Bw*R, C]
//     LSTM layer l  (site 16 + l)   e = (((z*T + t)*N + node) * L + unit            over [G*Bw, T, N, L]
//     head input    (site 32)       e = (z*N + node) * L + unit                     over [G*Bw*N, L]
// rng points at two device words {seed, pass counter}; wf_rng_advance bumps the counter after a forward+backward
// pair, on the stream, so a captured CUDA graph draws fresh masks on every replay.
#pragma once
#include <stdint.h>

#define WF_SITE_GCN 0
#define WF_SITE_LSTM 16
#define WF_SITE_HEAD 32

struct DropCfg {
  const unsigned long long* rng;  // device: {seed, pass counter}; nullptr = dropout off
  uint32_t thr;                   // drop when word < thr; thr = round(p * 2^32)
  float scale;                    // 1 / (1 - p)
  int site;
};

static inline DropCfg wf_drop_cfg(float p, const unsigned long long* rng, int site) {
  DropCfg d;
  d.rng = (p > 0.f && rng != nullptr) ? rng : nullptr;
  double t = (double)p * 4294967296.0;
  d.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)(t + 0.5);
  d.scale = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
  d.site = site;
  return d;
}

__device__ __forceinline__ uint4 wf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Keep factors (0 or 1/(1-p)) of the four elements 4*e4 .. 4*e4 + 3 of a site.
struct DropState { uint32_t k0, k1, c3, thr, site; float scale; };
__device__ __forceinline__ DropState wf_drop_state(const DropCfg& d) {
  DropState s;
  const unsigned long long seed = d.rng[0], pass = d.rng[1];
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(pass >> 32);
  s.c3 = (uint32_t)pass; s.thr = d.thr; s.site = (uint32_t)d.site; s.scale = d.scale;
  return s;
}
__device__ __forceinline__ void wf_drop4(const DropState& s, unsigned long long e4, float* m) {
  const uint4 r = wf_philox4x32_10((uint32_t)e4, (uint32_t)(e4 >> 32), s.site, s.c3, s.k0, s.k1);
  m[0] = r.x >= s.thr ? s.scale : 0.f;
  m[1] = r.y >= s.thr ? s.scale : 0.f;
  m[2] = r.z >= s.thr ? s.scale : 0.f;
  m[3] = r.w >= s.thr ? s.scale : 0.f;
}
