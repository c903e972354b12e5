// CUDA executor of the C++ engine: turns one negotiated Response (possibly a fused group of
// tensors) into kernel launches on the engine's communication stream (channel 0).
//
//   wait(ready events) -> batched pack into the symmetric staging buffer -> ONE collective
//   kernel over peer memory -> batched unpack -> record done event
//
// Arbitrary sizes are handled by chunking through the staging buffer.  The DistributedOptimizer
// hot path does not come through here: its gradients already live in symmetric memory.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "comm.h"
#include "gpu_work.h"

using namespace sdlc;

extern "C" {
int sdlc_allreduce(void*, int, int, long long, long long, int, int, float, float, int, unsigned long long);
int sdlc_pull(void*, int, int, void*, const long long*, const long long*, const long long*, long long, unsigned long long);
int sdlc_reducescatter(void*, int, int, void*, long long, long long, int, int, float, float, long long, unsigned long long);
int sdlc_pack(void*, int, const unsigned long long*, const long long*, const long long*, int, int, unsigned long long);
int sdlc_bcast_multimem(void*, int, int, const void*, void*, long long, int, unsigned long long);
}

namespace {

size_t esz_of(uint32_t dt) {
  switch (dt) {
    case kU8: case kI8: case kBool: case kF8: return 1;
    case kF16: case kBF16: return 2;
    case kI32: case kF32: return 4;
    default: return 8;
  }
}
long long align16(long long x) { return (x + 15) & ~15ll; }

struct Fail {
  char* err;
  int len;
  int operator()(Comm* c, const char* what) const {
    snprintf(err, len, "%s: %s", what, c->last_error.c_str());
    return -1;
  }
};

int exec_allreduce(Comm* c, const sdl::GpuWork* w, unsigned long long s, const Fail& fail) {
  const int buf = c->staging_buf[0];
  const long long cap = (long long)c->bufs[buf].bytes;
  const long long esz = (long long)esz_of(w->dtype);
  const float pre = (float)w->prescale, post = (float)w->postscale;
  uint32_t i = 0;
  while (i < w->n) {
    const long long bytes_i = w->numel[i] * esz;
    if (align16(bytes_i) > cap) {
      // one tensor larger than the staging buffer: stream it through in chunks
      const long long chunk_elems = (cap / 16) * (16 / (esz > 16 ? 16 : esz));
      for (long long e0 = 0; e0 < w->numel[i]; e0 += chunk_elems) {
        const long long ne = std::min(chunk_elems, w->numel[i] - e0);
        unsigned long long pin = w->in[i] ? (unsigned long long)(uintptr_t)w->in[i] + e0 * esz : 0;
        unsigned long long pout = w->out[i] ? (unsigned long long)(uintptr_t)w->out[i] + e0 * esz : 0;
        long long off = 0, nb = ne * esz;
        if (sdlc_pack(c, buf, &pin, &off, &nb, 1, 1, s)) return fail(c, "pack");
        if (sdlc_allreduce(c, 0, buf, 0, ne, (int)w->dtype, (int)w->op, pre, post, 0, s) < 0) return fail(c, "allreduce");
        if (sdlc_pack(c, buf, &pout, &off, &nb, 1, 0, s)) return fail(c, "unpack");
      }
      ++i;
      continue;
    }
    // greedy group of tensors that fits the staging buffer
    std::vector<unsigned long long> pin, pout;
    std::vector<long long> offs, sizes;
    long long total = 0;
    while (i < w->n && (long long)pin.size() < 1024) {
      const long long nb = w->numel[i] * esz;
      if (total + align16(nb) > cap) break;
      pin.push_back((unsigned long long)(uintptr_t)w->in[i]);
      pout.push_back((unsigned long long)(uintptr_t)w->out[i]);
      offs.push_back(total);
      sizes.push_back(nb);
      total += align16(nb);
      ++i;
    }
    // padding bytes between tensors hold stale data; they are reduced but never unpacked
    if (sdlc_pack(c, buf, pin.data(), offs.data(), sizes.data(), (int)pin.size(), 1, s)) return fail(c, "pack");
    if (sdlc_allreduce(c, 0, buf, 0, total / esz, (int)w->dtype, (int)w->op, pre, post, 0, s) < 0) return fail(c, "allreduce");
    if (sdlc_pack(c, buf, pout.data(), offs.data(), sizes.data(), (int)pout.size(), 0, s)) return fail(c, "unpack");
  }
  return 0;
}

int exec_broadcast(Comm* c, const sdl::GpuWork* w, unsigned long long s, const Fail& fail) {
  const int buf = c->staging_buf[0];
  const long long cap = (long long)c->bufs[buf].bytes & ~15ll;
  const long long total = w->numel[0] * (long long)esz_of(w->dtype);
  char* user = static_cast<char*>(w->out[0]);
  const bool mc = c->bufs[buf].mc != nullptr && c->use_mc && c->world > 1;
  for (long long b0 = 0; b0 < total || b0 == 0; b0 += cap) {
    const long long nb = std::min(cap, total - b0);
    if (mc && (nb & 15) == 0) {   // rank-independent choice: every rank must run the same kernel
      if (sdlc_bcast_multimem(c, 0, buf, user + b0, user + b0, nb, w->root, s)) return fail(c, "bcast_multimem");
    } else {
      if (c->rank == w->root) {
        unsigned long long p = (unsigned long long)(uintptr_t)(user + b0);
        long long off = 0, n = nb;
        if (sdlc_pack(c, buf, &p, &off, &n, 1, 1, s)) return fail(c, "pack");
      }
      long long so[kMaxPeers] = {0}, nn[kMaxPeers] = {0}, dof[kMaxPeers] = {0};
      if (c->rank != w->root) nn[w->root] = nb;
      if (sdlc_pull(c, 0, buf, user + b0, so, nn, dof, nb, s)) return fail(c, "pull");
    }
    if (total == 0) break;
  }
  return 0;
}

int exec_allgather(Comm* c, const sdl::GpuWork* w, unsigned long long s, const Fail& fail) {
  const int buf = c->staging_buf[0];
  const long long cap = (long long)c->bufs[buf].bytes & ~15ll;
  const long long row = w->row_elems * (long long)esz_of(w->dtype);
  long long bytes[kMaxPeers] = {0}, prefix[kMaxPeers] = {0}, maxb = 0, acc = 0;
  for (int p = 0; p < c->world; ++p) {
    bytes[p] = w->dim0[p] * row;
    prefix[p] = acc;
    acc += bytes[p];
    maxb = std::max(maxb, bytes[p]);
  }
  for (long long b0 = 0; b0 < maxb || b0 == 0; b0 += cap) {
    const long long mine = std::max(0ll, std::min(cap, bytes[c->rank] - b0));
    if (mine > 0) {
      unsigned long long p = (unsigned long long)(uintptr_t)(static_cast<char*>(w->in[0]) + b0);
      long long off = 0, n = mine;
      if (sdlc_pack(c, buf, &p, &off, &n, 1, 1, s)) return fail(c, "pack");
    }
    long long so[kMaxPeers] = {0}, nn[kMaxPeers] = {0}, dof[kMaxPeers] = {0}, work = 0;
    for (int p = 0; p < c->world; ++p) {
      nn[p] = std::max(0ll, std::min(cap, bytes[p] - b0));
      dof[p] = prefix[p] + b0;
      work += nn[p];
    }
    if (sdlc_pull(c, 0, buf, w->out[0], so, nn, dof, work, s)) return fail(c, "pull");
    if (maxb == 0) break;
  }
  return 0;
}

int exec_alltoall(Comm* c, const sdl::GpuWork* w, unsigned long long s, const Fail& fail) {
  const int buf = c->staging_buf[0], W = c->world, st = w->splits_stride;
  const long long row = w->row_elems * (long long)esz_of(w->dtype);
  const long long segcap = (((long long)c->bufs[buf].bytes / W) & ~15ll);
  long long send_prefix[kMaxPeers] = {0}, recv_prefix[kMaxPeers] = {0}, maxseg = 0, a = 0, b = 0;
  for (int d = 0; d < W; ++d) {
    send_prefix[d] = a;
    a += (long long)w->splits[c->rank * st + d] * row;
    recv_prefix[d] = b;
    b += (long long)w->splits[d * st + c->rank] * row;
    for (int q = 0; q < W; ++q) maxseg = std::max(maxseg, (long long)w->splits[q * st + d] * row);
  }
  for (long long b0 = 0; b0 < maxseg || b0 == 0; b0 += segcap) {
    unsigned long long ptrs[kMaxPeers];
    long long offs[kMaxPeers], sizes[kMaxPeers];
    int n = 0;
    for (int d = 0; d < W; ++d) {
      const long long seg = (long long)w->splits[c->rank * st + d] * row;
      const long long nb = std::max(0ll, std::min(segcap, seg - b0));
      if (nb > 0) {
        ptrs[n] = (unsigned long long)(uintptr_t)(static_cast<char*>(w->in[0]) + send_prefix[d] + b0);
        offs[n] = d * segcap;
        sizes[n] = nb;
        ++n;
      }
    }
    if (n > 0 && sdlc_pack(c, buf, ptrs, offs, sizes, n, 1, s)) return fail(c, "pack");
    long long so[kMaxPeers] = {0}, nn[kMaxPeers] = {0}, dof[kMaxPeers] = {0};
    for (int p = 0; p < W; ++p) {
      const long long seg = (long long)w->splits[p * st + c->rank] * row;
      so[p] = c->rank * segcap;
      nn[p] = std::max(0ll, std::min(segcap, seg - b0));
      dof[p] = recv_prefix[p] + b0;
    }
    // grid bound = the largest receive volume of any rank in this round (same value everywhere)
    long long work = 0;
    for (int d = 0; d < W; ++d) {
      long long recv_d = 0;
      for (int p = 0; p < W; ++p) recv_d += std::max(0ll, std::min(segcap, (long long)w->splits[p * st + d] * row - b0));
      work = std::max(work, recv_d);
    }
    if (sdlc_pull(c, 0, buf, w->out[0], so, nn, dof, work, s)) return fail(c, "pull");
    if (maxseg == 0) break;
  }
  return 0;
}

int exec_reducescatter(Comm* c, const sdl::GpuWork* w, unsigned long long s, const Fail& fail) {
  const int buf = c->staging_buf[0], W = c->world;
  const long long esz = (long long)esz_of(w->dtype);
  const long long cap_elems = (((long long)c->bufs[buf].bytes & ~15ll) / esz);
  const long long d0 = w->dim0[0], row = w->row_elems, total = w->numel[0];
  long long my_lo = 0;
  for (int p = 0; p < c->rank; ++p) my_lo += (d0 / W + (p < d0 % W ? 1 : 0)) * row;
  const long long my_cnt = (d0 / W + (c->rank < d0 % W ? 1 : 0)) * row;
  for (long long e0 = 0; e0 < total || e0 == 0; e0 += cap_elems) {
    const long long ne = std::min(cap_elems, total - e0);
    unsigned long long p = (unsigned long long)(uintptr_t)(static_cast<char*>(w->in[0]) + e0 * esz);
    long long off = 0, nb = ne * esz;
    if (nb > 0 && sdlc_pack(c, buf, &p, &off, &nb, 1, 1, s)) return fail(c, "pack");
    const long long lo = std::max(my_lo, e0), hi = std::min(my_lo + my_cnt, e0 + ne);
    const long long cnt = std::max(0ll, hi - lo);
    char* out = static_cast<char*>(w->out[0]) + (cnt > 0 ? (lo - my_lo) * esz : 0);
    if (sdlc_reducescatter(c, 0, buf, out, cnt > 0 ? lo - e0 : 0, cnt, (int)w->dtype, (int)w->op, (float)w->prescale,
                           (float)w->postscale, (ne + W - 1) / W, s))
      return fail(c, "reducescatter");
    if (total == 0) break;
  }
  return 0;
}

}  // namespace

extern "C" int sdlc_engine_exec(void* ctx, const sdl::GpuWork* w, uint64_t* done_event, char* err, int errlen) {
  Comm* c = static_cast<Comm*>(ctx);
  cudaSetDevice(c->device);
  if (c->staging_buf[0] < 0 || c->pad_buf < 0) {
    snprintf(err, errlen, "CUDA communicator is not fully initialised");
    return -1;
  }
  cudaStream_t stream = c->streams[0];
  const unsigned long long s = reinterpret_cast<unsigned long long>(stream);
  for (uint32_t i = 0; i < w->n; ++i)
    if (w->ready_events[i]) {
      cudaEvent_t re = reinterpret_cast<cudaEvent_t>(w->ready_events[i]);
      cudaStreamWaitEvent(stream, re, 0);
      c->release_event(re);   // the wait is issued: the producer-side event can be recycled
    }
  Fail fail{err, errlen};
  int rc;
  switch (w->type) {
    case 0: rc = exec_allreduce(c, w, s, fail); break;
    case 1: rc = exec_allgather(c, w, s, fail); break;
    case 2: rc = exec_broadcast(c, w, s, fail); break;
    case 3: rc = exec_alltoall(c, w, s, fail); break;
    case 4: rc = exec_reducescatter(c, w, s, fail); break;
    default: snprintf(err, errlen, "unsupported CUDA collective type %u", w->type); rc = -1;
  }
  if (rc != 0) return rc;
  // one completion event shared by the w->n tensors of this (fused) response; each handle releases it once
  cudaEvent_t ev = c->acquire_event((int)w->n);
  if (!ev || cudaEventRecord(ev, stream) != cudaSuccess) {
    snprintf(err, errlen, "cudaEventRecord failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  *done_event = reinterpret_cast<uint64_t>(ev);
  return 0;
}
