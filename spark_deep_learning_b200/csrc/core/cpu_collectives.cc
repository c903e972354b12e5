#include "cpu_collectives.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace sdl {

// ---- half / bfloat16 <-> float on the host -------------------------------------------------
static inline float bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1);                                             // RNE
  return (uint16_t)(u >> 16);
}
static inline float f16_to_f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff, u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else {
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400));
      u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7f800000u | (man << 13);
  } else {
    u = sign | ((exp + 112) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t sign = (u >> 16) & 0x8000;
  int32_t exp = (int32_t)((u >> 23) & 0xff) - 127 + 15;
  uint32_t man = u & 0x7fffff;
  if (((u >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00 | (man ? 0x200 : 0));
  if (exp >= 31) return (uint16_t)(sign | 0x7c00);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    man |= 0x800000;
    uint32_t shift = (uint32_t)(14 - exp);
    uint32_t h = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
  uint32_t rem = man & 0x1fff;
  if (rem > 0x1000 || (rem == 0x1000 && (h & 1))) ++h;
  return (uint16_t)(sign | h);
}

template <typename T>
static inline T apply(T a, T b, uint32_t op) {
  switch (op) {
    case kMin: return std::min(a, b);
    case kMax: return std::max(a, b);
    case kProduct: return (T)(a * b);
    default: return (T)(a + b);
  }
}

template <typename T>
static void reduce_typed(T* d, const T* s, int64_t n, uint32_t op) {
  switch (op) {
    case kMin: for (int64_t i = 0; i < n; ++i) d[i] = std::min(d[i], s[i]); break;
    case kMax: for (int64_t i = 0; i < n; ++i) d[i] = std::max(d[i], s[i]); break;
    case kProduct: for (int64_t i = 0; i < n; ++i) d[i] = (T)(d[i] * s[i]); break;
    default: for (int64_t i = 0; i < n; ++i) d[i] = (T)(d[i] + s[i]); break;
  }
}

void reduce_into(void* dst, const void* src, int64_t n, uint32_t dtype, uint32_t op) {
  switch (dtype) {
    case kF32: reduce_typed((float*)dst, (const float*)src, n, op); break;
    case kF64: reduce_typed((double*)dst, (const double*)src, n, op); break;
    case kI32: reduce_typed((int32_t*)dst, (const int32_t*)src, n, op); break;
    case kI64: reduce_typed((int64_t*)dst, (const int64_t*)src, n, op); break;
    case kU8: reduce_typed((uint8_t*)dst, (const uint8_t*)src, n, op); break;
    case kI8: reduce_typed((int8_t*)dst, (const int8_t*)src, n, op); break;
    case kBool: {
      uint8_t* d = (uint8_t*)dst;
      const uint8_t* s = (const uint8_t*)src;
      if (op == kMin || op == kProduct)
        for (int64_t i = 0; i < n; ++i) d[i] = (uint8_t)(d[i] && s[i]);
      else
        for (int64_t i = 0; i < n; ++i) d[i] = (uint8_t)(d[i] || s[i]);
      break;
    }
    case kBF16: {
      uint16_t* d = (uint16_t*)dst;
      const uint16_t* s = (const uint16_t*)src;
      for (int64_t i = 0; i < n; ++i) d[i] = f_to_bf16(apply(bf16_to_f(d[i]), bf16_to_f(s[i]), op));
      break;
    }
    case kF16: {
      uint16_t* d = (uint16_t*)dst;
      const uint16_t* s = (const uint16_t*)src;
      for (int64_t i = 0; i < n; ++i) d[i] = f_to_f16(apply(f16_to_f(d[i]), f16_to_f(s[i]), op));
      break;
    }
    default: break;
  }
}

void scale_buffer(void* buf, int64_t n, uint32_t dtype, double factor) {
  if (factor == 1.0) return;
  switch (dtype) {
    case kF32: { float* p = (float*)buf; float f = (float)factor; for (int64_t i = 0; i < n; ++i) p[i] *= f; break; }
    case kF64: { double* p = (double*)buf; for (int64_t i = 0; i < n; ++i) p[i] *= factor; break; }
    case kBF16: { uint16_t* p = (uint16_t*)buf; float f = (float)factor; for (int64_t i = 0; i < n; ++i) p[i] = f_to_bf16(bf16_to_f(p[i]) * f); break; }
    case kF16: { uint16_t* p = (uint16_t*)buf; float f = (float)factor; for (int64_t i = 0; i < n; ++i) p[i] = f_to_f16(f16_to_f(p[i]) * f); break; }
    case kI32: { int32_t* p = (int32_t*)buf; for (int64_t i = 0; i < n; ++i) p[i] = (int32_t)std::llround(p[i] * factor); break; }
    case kI64: { int64_t* p = (int64_t*)buf; for (int64_t i = 0; i < n; ++i) p[i] = (int64_t)std::llround((double)p[i] * factor); break; }
    default: break;
  }
}

bool CpuCollectives::allreduce(void* buf, int64_t numel, uint32_t dtype, uint32_t op) {
  const int world = shm_->world(), rank = shm_->rank();
  if (world == 1 || numel == 0) return true;
  const int64_t esz = (int64_t)dtype_size(dtype);
  const int64_t chunk = (int64_t)(shm_->stage_bytes() / esz);
  uint8_t* b = (uint8_t*)buf;
  for (int64_t off = 0; off < numel; off += chunk) {
    const int64_t n = std::min(chunk, numel - off);
    memcpy(shm_->stage(rank), b + off * esz, n * esz);
    if (!shm_->barrier(0)) return false;
    const int64_t per = (n + world - 1) / world;
    const int64_t lo = std::min(n, rank * per), hi = std::min(n, lo + per);
    if (hi > lo) {
      uint8_t* r = shm_->result() + lo * esz;
      memcpy(r, shm_->stage(0) + lo * esz, (hi - lo) * esz);
      for (int p = 1; p < world; ++p) reduce_into(r, shm_->stage(p) + lo * esz, hi - lo, dtype, op);
    }
    if (!shm_->barrier(0)) return false;
    memcpy(b + off * esz, shm_->result(), n * esz);
    if (!shm_->barrier(0)) return false;
  }
  return true;
}

bool CpuCollectives::reducescatter(const void* in, void* out, const int64_t* numel_per_rank,
                                   uint32_t dtype, uint32_t op) {
  const int world = shm_->world(), rank = shm_->rank();
  const int64_t esz = (int64_t)dtype_size(dtype);
  int64_t total = 0, my_lo = 0;
  for (int p = 0; p < world; ++p) {
    if (p == rank) my_lo = total;
    total += numel_per_rank[p];
  }
  const int64_t my_hi = my_lo + numel_per_rank[rank];
  if (world == 1) {
    memcpy(out, in, total * esz);
    return true;
  }
  const int64_t chunk = (int64_t)(shm_->stage_bytes() / esz);
  const uint8_t* b = (const uint8_t*)in;
  for (int64_t off = 0; off < total; off += chunk) {
    const int64_t n = std::min(chunk, total - off);
    memcpy(shm_->stage(rank), b + off * esz, n * esz);
    if (!shm_->barrier(0)) return false;
    const int64_t per = (n + world - 1) / world;
    const int64_t lo = std::min(n, rank * per), hi = std::min(n, lo + per);
    if (hi > lo) {
      uint8_t* r = shm_->result() + lo * esz;
      memcpy(r, shm_->stage(0) + lo * esz, (hi - lo) * esz);
      for (int p = 1; p < world; ++p) reduce_into(r, shm_->stage(p) + lo * esz, hi - lo, dtype, op);
    }
    if (!shm_->barrier(0)) return false;
    const int64_t a = std::max(off, my_lo), z = std::min(off + n, my_hi);
    if (z > a) memcpy((uint8_t*)out + (a - my_lo) * esz, shm_->result() + (a - off) * esz, (z - a) * esz);
    if (!shm_->barrier(0)) return false;
  }
  return true;
}

bool CpuCollectives::allgather(const void* in, void* out, const int64_t* bytes_per_rank) {
  const int world = shm_->world(), rank = shm_->rank();
  int64_t maxb = 0;
  for (int p = 0; p < world; ++p) maxb = std::max(maxb, bytes_per_rank[p]);
  if (world == 1) {
    memcpy(out, in, bytes_per_rank[0]);
    return true;
  }
  const int64_t stage = (int64_t)shm_->stage_bytes();
  for (int64_t off = 0; off < maxb; off += stage) {
    const int64_t mine = std::max<int64_t>(0, std::min(stage, bytes_per_rank[rank] - off));
    if (mine > 0) memcpy(shm_->stage(rank), (const uint8_t*)in + off, mine);
    if (!shm_->barrier(0)) return false;
    int64_t prefix = 0;
    for (int p = 0; p < world; ++p) {
      const int64_t n = std::max<int64_t>(0, std::min(stage, bytes_per_rank[p] - off));
      if (n > 0) memcpy((uint8_t*)out + prefix + off, shm_->stage(p), n);
      prefix += bytes_per_rank[p];
    }
    if (!shm_->barrier(0)) return false;
  }
  return true;
}

bool CpuCollectives::broadcast(void* buf, int64_t bytes, int root) {
  const int world = shm_->world(), rank = shm_->rank();
  if (world == 1) return true;
  const int64_t stage = (int64_t)shm_->stage_bytes();
  for (int64_t off = 0; off < bytes; off += stage) {
    const int64_t n = std::min(stage, bytes - off);
    if (rank == root) memcpy(shm_->result(), (uint8_t*)buf + off, n);
    if (!shm_->barrier(0)) return false;
    if (rank != root) memcpy((uint8_t*)buf + off, shm_->result(), n);
    if (!shm_->barrier(0)) return false;
  }
  return true;
}

bool CpuCollectives::alltoall(const void* in, void* out, const int64_t* m, const int64_t*) {
  // m[src * world + dst] = bytes src sends to dst (identical on every rank)
  const int world = shm_->world(), rank = shm_->rank();
  const int64_t stage = (int64_t)shm_->stage_bytes();
  int64_t send_off = 0;
  for (int d = 0; d < world; ++d) {
    int64_t maxb = 0;
    for (int p = 0; p < world; ++p) maxb = std::max(maxb, m[p * world + d]);
    const int64_t mine_total = m[rank * world + d];
    if (world == 1) {
      memcpy(out, in, mine_total);
      return true;
    }
    for (int64_t off = 0; off < maxb; off += stage) {
      const int64_t mine = std::max<int64_t>(0, std::min(stage, mine_total - off));
      if (mine > 0) memcpy(shm_->stage(rank), (const uint8_t*)in + send_off + off, mine);
      if (!shm_->barrier(0)) return false;
      if (rank == d) {
        int64_t prefix = 0;
        for (int p = 0; p < world; ++p) {
          const int64_t n = std::max<int64_t>(0, std::min(stage, m[p * world + d] - off));
          if (n > 0) memcpy((uint8_t*)out + prefix + off, shm_->stage(p), n);
          prefix += m[p * world + d];
        }
      }
      if (!shm_->barrier(0)) return false;
    }
    send_off += mine_total;
  }
  return true;
}

}  // namespace sdl
