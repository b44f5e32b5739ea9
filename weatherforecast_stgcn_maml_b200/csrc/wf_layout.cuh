// Offsets of the trainable flat parameter buffer (state_dict order) and of the
// pre-transposed weight buffer used by the tensor-core backward products.
#pragma once

struct LstmLayout {
  long long w_ih[8], w_hh[8], b_ih[8], b_hh[8], head_w, head_b, total;
  // transposed buffer: whhT[l] = W_hh[l]^T [L, 4L]; wihT[l] = W_ih[l]^T [L, 4L] for l >= 1
  long long whhT[8], wihT[8], totalT;
};

static inline LstmLayout lstm_layout(int layers, int F, int L, int O) {
  LstmLayout p;
  long long off = 0;
  for (int l = 0; l < layers; ++l) {
    int kin = l == 0 ? F : L;
    p.w_ih[l] = off; off += 4LL * L * kin;
    p.w_hh[l] = off; off += 4LL * L * L;
    p.b_ih[l] = off; off += 4LL * L;
    p.b_hh[l] = off; off += 4LL * L;
  }
  p.head_w = off; off += (long long)O * L;
  p.head_b = off; off += O;
  p.total = off;
  long long t = 0;
  for (int l = 0; l < layers; ++l) {
    p.whhT[l] = t; t += 4LL * L * L;
    p.wihT[l] = l > 0 ? t : -1;
    if (l > 0) t += 4LL * L * L;
  }
  p.totalT = t;
  return p;
}
