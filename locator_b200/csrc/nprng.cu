// numpy's legacy random stream on the host, for the draws that bound the jacknife sweep.
//
// The reference draws np.random.binomial(2, af[site], n_pred) for every removed site of every jacknife
// replicate (locator/locator.py:722-727: 2.5 M scalar-p draws per replicate at config 5) and
// binomial(2, af) per missing call in replace_md (:258-261), all from numpy's global RandomState.  To
// reproduce the reference's indices the stream has to be consumed in exactly that order; numpy spends
// ~50 ns per draw on broadcasting overhead.  This file restates the generator so that the same draws cost
// a few ns each: MT19937 (Matsumoto & Nishimura 1998: state of 624 words, the published recurrence and
// tempering), numpy's 53-bit double from two outputs ((a >> 5) * 2^26 + (b >> 6)) / 2^53, and the
// inversion sampler numpy's legacy binomial uses while min(p, 1 - p) * n <= 30 (sequential search from
// X = 0 with px = q^n, restart above the bound min(n, np + 10 sqrt(npq + 1))).  The caller copies the
// state out of / back into np.random (get_state / set_state).  Pure host code; exp / log / sqrt are the
// C library's, as in numpy.  tests/test_host.py checks draws and the stream position against numpy.
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kN = 624, kM = 397;

struct Mt {
  uint32_t* key;
  int pos;
};

inline void mt_refill(uint32_t* k) {
  auto twist = [](uint32_t u, uint32_t v) -> uint32_t {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  };
  int i = 0;
  for (; i < kN - kM; ++i) k[i] = k[i + kM] ^ twist(k[i], k[i + 1]);
  for (; i < kN - 1; ++i) k[i] = k[i + (kM - kN)] ^ twist(k[i], k[i + 1]);
  k[kN - 1] = k[kM - 1] ^ twist(k[kN - 1], k[0]);
}

inline uint32_t mt_next(Mt& s) {
  if (s.pos >= kN) {
    mt_refill(s.key);
    s.pos = 0;
  }
  uint32_t y = s.key[s.pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

inline double mt_double(Mt& s) {
  const uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

// parameters of the inversion sampler for one (n, p), p <= 0.5 after the caller's reflection
struct Inv {
  double p, q, qn;
  int64_t bound;
};

inline Inv inv_setup(int64_t n, double p) {
  Inv v;
  v.p = p;
  v.q = 1.0 - p;
  v.qn = exp((double)n * log(v.q));
  const double np = (double)n * p;
  const double b = np + 10.0 * sqrt(np * v.q + 1);
  v.bound = (int64_t)((double)n < b ? (double)n : b);
  return v;
}

inline int64_t inv_draw(Mt& s, int64_t n, const Inv& v) {
  int64_t X = 0;
  double px = v.qn;
  double U = mt_double(s);
  while (U > px) {
    X++;
    if (X > v.bound) {
      X = 0;
      px = v.qn;
      U = mt_double(s);
    } else {
      U -= px;
      px = ((double)(n - X + 1) * v.p * px) / ((double)X * v.q);
    }
  }
  return X;
}

// The same sampler with the success probabilities px(X) tabulated once per (n, p): the recurrence is
// evaluated with the same operations in the same order, so the doubles -- and every comparison -- are those
// of inv_draw, without a division per step.
constexpr int kTabN = 16;
struct InvTab {
  double P[kTabN + 1];
  int64_t bound;
};

inline InvTab tab_setup(int64_t n, const Inv& v) {
  InvTab t;
  t.bound = v.bound;
  t.P[0] = v.qn;
  for (int64_t X = 1; X <= v.bound && X <= kTabN; ++X)
    t.P[X] = ((double)(n - X + 1) * v.p * t.P[X - 1]) / ((double)X * v.q);
  return t;
}

inline int64_t tab_draw(Mt& s, const InvTab& t) {
  int64_t X = 0;
  double U = mt_double(s);
  while (U > t.P[X]) {
    X++;
    if (X > t.bound) {
      X = 0;
      U = mt_double(s);
    } else {
      U -= t.P[X - 1];
    }
  }
  return X;
}

// n = 2 (every draw of the reference): the bound is always 2, so the search is three compares.  Written
// without data-dependent branches (the genotype is as unpredictable as U); the subtractions are the
// sampler's own, so every comparison sees the same doubles.  Only the restart (U beyond the rounded
// total mass, practically never) branches.
inline int64_t tab_draw2(Mt& s, const InvTab& t) {
  for (;;) {
    const double U = mt_double(s);
    const double U1 = U - t.P[0], U2 = U1 - t.P[1];
    const int c0 = U > t.P[0];
    const int c1 = c0 & (U1 > t.P[1]);
    const int c2 = c1 & (U2 > t.P[2]);
    if (!c2) return c0 + c1;
  }
}

// numpy's bounded integer for the legacy shuffle: rejection below the smallest all-ones mask >= max
// (32-bit outputs while max fits, two outputs -- high word first -- beyond)
inline uint64_t mt_interval(Mt& s, uint64_t max) {
  if (max == 0) return 0;
  uint64_t mask = max;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  mask |= mask >> 32;
  uint64_t v;
  if (max <= 0xffffffffull) {
    while ((v = (mt_next(s) & mask)) > max) {
    }
  } else {
    for (;;) {
      const uint64_t hi = mt_next(s), lo = mt_next(s);
      v = ((hi << 32) | lo) & mask;
      if (v <= max) break;
    }
  }
  return v;
}

}  // namespace

extern "C" {

// RandomState.permutation(n) (and with it choice(n, size, replace=False) = permutation(n)[:size],
// locator.py:299, :722): arange(n) shuffled by numpy's legacy Fisher-Yates -- for i = n-1 .. 1 swap element i
// with element random_interval(i) -- from the same MT19937 state, advanced in place.
int loc_np_legacy_permutation(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int64_t* out) {
  LOC_CHECK(mt_key != nullptr && mt_pos != nullptr && n >= 0 && (n == 0 || out != nullptr),
            "loc_np_legacy_permutation: bad arguments");
  LOC_CHECK(*mt_pos >= 0 && *mt_pos <= kN, "loc_np_legacy_permutation: bad stream position");
  Mt s{mt_key, (int)*mt_pos};
  for (int64_t i = 0; i < n; ++i) out[i] = i;
  for (int64_t i = n - 1; i >= 1; --i) {
    const int64_t j = (int64_t)mt_interval(s, (uint64_t)i);
    const int64_t t = out[i];
    out[i] = out[j];
    out[j] = t;
  }
  *mt_pos = s.pos;
  return 0;
}

// out[i * reps + r] = the r-th of `reps` consecutive RandomState.binomial(n, p[i]) draws, sites in order.
// mt_key[624] / *mt_pos: the MT19937 state of np.random.get_state(), advanced in place.
// Returns 0; 1 when some (n, p[i]) would take numpy's BTPE branch (min(p, 1-p) * n > 30) or n > 255 --
// nothing has been drawn then and the caller uses numpy itself; 2 for p outside [0, 1] or NaN.
int loc_np_legacy_binomial(uint32_t* mt_key, int32_t* mt_pos, int64_t n, const double* p, int64_t n_p, int64_t reps,
                           uint8_t* out) {
  LOC_CHECK(mt_key != nullptr && mt_pos != nullptr && n >= 0 && n_p >= 0 && reps >= 0, "loc_np_legacy_binomial: bad arguments");
  LOC_CHECK(n_p == 0 || reps == 0 || (p != nullptr && out != nullptr), "loc_np_legacy_binomial: null array");
  LOC_CHECK(*mt_pos >= 0 && *mt_pos <= kN, "loc_np_legacy_binomial: bad stream position");
  if (n > 255) return 1;
  for (int64_t i = 0; i < n_p; ++i) {
    const double pi = p[i];
    if (!(pi >= 0.0 && pi <= 1.0)) {
      loc::fail("loc_np_legacy_binomial: p outside [0, 1]", __FILE__, __LINE__);
      return 2;
    }
    const double small = pi <= 0.5 ? pi : 1.0 - pi;
    if (small * (double)n > 30.0) return 1;
  }
  Mt s{mt_key, (int)*mt_pos};
  for (int64_t i = 0; i < n_p; ++i) {
    const double pi = p[i];
    const bool reflect = !(pi <= 0.5);
    const Inv v = inv_setup(n, reflect ? 1.0 - pi : pi);
    uint8_t* o = out + i * reps;
    if (n == 2 && v.bound == 2) {
      const InvTab t = tab_setup(n, v);
      if (reflect) {
        for (int64_t r = 0; r < reps; ++r) o[r] = (uint8_t)(2 - tab_draw2(s, t));
      } else {
        for (int64_t r = 0; r < reps; ++r) o[r] = (uint8_t)tab_draw2(s, t);
      }
    } else if (n <= kTabN) {
      const InvTab t = tab_setup(n, v);
      if (reflect) {
        for (int64_t r = 0; r < reps; ++r) o[r] = (uint8_t)(n - tab_draw(s, t));
      } else {
        for (int64_t r = 0; r < reps; ++r) o[r] = (uint8_t)tab_draw(s, t);
      }
    } else if (reflect) {
      for (int64_t r = 0; r < reps; ++r) o[r] = (uint8_t)(n - inv_draw(s, n, v));
    } else {
      for (int64_t r = 0; r < reps; ++r) o[r] = (uint8_t)inv_draw(s, n, v);
    }
  }
  *mt_pos = s.pos;
  return 0;
}

}  // extern "C"
