#include "layout.h"

#include <algorithm>
#include <cstdlib>

namespace hptb {

hptb_status broadcast_shape(const int64_t* a, int na, const int64_t* b, int nb, int64_t* out, int* nout) {
  int n = std::max(na, nb);
  if (n > HPTB_MAX_DIMS) return fail(HPTB_ERR_INVALID, "broadcast: ndim %d exceeds %d", n, HPTB_MAX_DIMS);
  for (int i = 0; i < n; ++i) {
    int ia = i - (n - na), ib = i - (n - nb);
    int64_t da = ia >= 0 ? a[ia] : 1, db = ib >= 0 ? b[ib] : 1;
    if (da == db || db == 1) out[i] = da;
    else if (da == 1) out[i] = db;
    else {
      // message format of predict_broadcast_shape (hpt-common/src/shape/shape_utils.rs:389-396), pinned by
      // hpt-tests/src/hpt_common/layout.rs:31-40
      auto fmt = [](const int64_t* s, int k) {
        std::string r = "[";
        for (int j = 0; j < k; ++j) r += (j ? ", " : "") + std::to_string((long long)s[j]);
        return r + "]";
      };
      return fail(HPTB_ERR_SHAPE, "Broadcasting error: broadcast failed at index %d, lhs shape: %s, rhs shape: %s", i,
                  fmt(a, na).c_str(), fmt(b, nb).c_str());
    }
  }
  *nout = n;
  return HPTB_OK;
}

hptb_status broadcast_strides(const hptb_tensor& t, const int64_t* shape, int ndim, int64_t* strides_out) {
  if (t.ndim > ndim) return fail(HPTB_ERR_SHAPE, "operand ndim %d exceeds result ndim %d", t.ndim, ndim);
  int off = ndim - t.ndim;
  for (int i = 0; i < ndim; ++i) {
    if (i < off) { strides_out[i] = 0; continue; }
    int64_t d = t.shape[i - off];
    if (d == shape[i]) strides_out[i] = (d == 1) ? 0 : t.strides[i - off];
    else if (d == 1) strides_out[i] = 0;
    else
      return fail(HPTB_ERR_SHAPE, "cannot broadcast dim %d of size %lld to %lld", i - off, (long long)d,
                  (long long)shape[i]);
  }
  return HPTB_OK;
}

void collapse(int ndim, const int64_t* shape, int nops, const int64_t (*strides)[HPTB_MAX_DIMS],
              const uint8_t* reduced, Collapsed* out) {
  Collapsed c;
  c.nops = nops;
  // 1. drop size-1 dims
  int idx[HPTB_MAX_DIMS];
  int n = 0;
  c.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    c.numel *= shape[i];
    if (shape[i] != 1) idx[n++] = i;
  }
  // 2. order: kept first (by |out stride| desc), then reduced (by |in stride| desc); stable
  auto key_op = [&](int d) { return (reduced && reduced[d]) ? (nops > 1 ? 1 : 0) : 0; };
  std::stable_sort(idx, idx + n, [&](int x, int y) {
    bool rx = reduced && reduced[x], ry = reduced && reduced[y];
    if (rx != ry) return !rx;  // kept before reduced
    int64_t sx = std::llabs(strides[key_op(x)][x]), sy = std::llabs(strides[key_op(y)][y]);
    return sx > sy;
  });
  // 3. merge
  int m = 0;
  for (int k = 0; k < n; ++k) {
    int d = idx[k];
    bool red = reduced && reduced[d];
    if (m > 0 && (bool)c.reduced[m - 1] == red) {
      bool ok = true;
      for (int o = 0; o < nops; ++o)
        if (c.strides[o][m - 1] != strides[o][d] * shape[d]) { ok = false; break; }
      if (ok) {
        c.shape[m - 1] *= shape[d];
        for (int o = 0; o < nops; ++o) c.strides[o][m - 1] = strides[o][d];
        continue;
      }
    }
    c.shape[m] = shape[d];
    c.reduced[m] = red;
    for (int o = 0; o < nops; ++o) c.strides[o][m] = strides[o][d];
    ++m;
  }
  c.ndim = m;
  // 4. classify (meaningful for elementwise use)
  if (m == 0) {
    c.launch_class = HPTB_CLASS_CONTIGUOUS;
  } else {
    bool inner_ok = c.strides[0][m - 1] == 1;
    for (int o = 1; o < nops; ++o)
      if (c.strides[o][m - 1] != 1 && c.strides[o][m - 1] != 0) inner_ok = false;
    if (inner_ok && m == 1) c.launch_class = HPTB_CLASS_CONTIGUOUS;
    else if (inner_ok) c.launch_class = HPTB_CLASS_INNER_CONTIGUOUS;
    else c.launch_class = HPTB_CLASS_STRIDED;
  }
  *out = c;
}

}  // namespace hptb
