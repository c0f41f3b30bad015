// stream.cu -- typed parallel_for fast paths: the five kernels of benchmarks/stream
// (benchmarks/stream/stream-kokkos.cpp:55-77: init/set, copy, scale, add, triad), instantiations of
// kb200/impl/ForKernel.hpp with 32-byte streaming loads and stores.
//
// Arithmetic is written without FMA contraction (__dadd_rn/__dmul_rn) so triad is bit-identical to
// the reference's OpenMP build compiled with -ffp-contract=off.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/impl/ForKernel.hpp>
#include <kb200/impl/ContigBody.hpp>

using namespace kb200;
using namespace kb200::Impl;

namespace {
// all arrays of one call must share the same 32-byte phase; otherwise the call falls back to 8-byte units

template <int NIN, class Op, int VBYTES>
struct StreamBody {
  static constexpr int E = VBYTES / 8;
  struct packet { RawVec<VBYTES> in[NIN > 0 ? NIN : 1]; };
  const double* in[NIN > 0 ? NIN : 1];
  double* out;
  double s;
  int64 head, nvec, tail;
  int aliased;  // an input is also the output: use coherent loads instead of the read-only path
  KB200_DEVICE_FUNCTION packet load(int64 u) const {
    packet p;
#pragma unroll
    for (int a = 0; a < NIN; ++a)
      p.in[a] = aliased ? ld_plain<VBYTES>(in[a] + head + u * E) : ld_stream<VBYTES>(in[a] + head + u * E);
    return p;
  }
  KB200_DEVICE_FUNCTION void store(const packet& p, int64 u) const {
    double v[NIN > 0 ? NIN : 1][E];
#pragma unroll
    for (int a = 0; a < NIN; ++a) memcpy(v[a], p.in[a].w, VBYTES);
    double r[E];
#pragma unroll
    for (int k = 0; k < E; ++k) r[k] = Op::apply(s, NIN > 0 ? v[0][k] : 0.0, NIN > 1 ? v[NIN > 1 ? 1 : 0][k] : 0.0);
    RawVec<VBYTES> o;
    memcpy(o.w, r, VBYTES);
    st_stream<VBYTES>(out + head + u * E, o);
  }
  KB200_FUNCTION int64 edge_count() const { return head + tail; }
  KB200_DEVICE_FUNCTION void edge(int64 k) const {
    const int64 i = k < head ? k : head + nvec * E + (k - head);
    out[i] = Op::apply(s, NIN > 0 ? in[0][i] : 0.0, NIN > 1 ? in[NIN > 1 ? 1 : 0][i] : 0.0);
  }
};

struct SetOp { KB200_DEVICE_FUNCTION static double apply(double s, double, double) { return s; } };
struct CopyOp { KB200_DEVICE_FUNCTION static double apply(double, double a, double) { return a; } };
struct ScaleOp { KB200_DEVICE_FUNCTION static double apply(double s, double c, double) { return __dmul_rn(s, c); } };
struct AddOp { KB200_DEVICE_FUNCTION static double apply(double, double a, double b) { return __dadd_rn(a, b); } };
struct TriadOp { KB200_DEVICE_FUNCTION static double apply(double s, double b, double c) { return __dadd_rn(b, __dmul_rn(s, c)); } };

template <int NIN, class Op, int VBYTES, int BLOCK, int UNROLL>
int launch(b200_instance* I, const double* i0, const double* i1, double* out, double s, int64 n, int bps) {
  StreamBody<NIN, Op, VBYTES> b;
  b.in[0] = i0;
  if (NIN > 1) b.in[NIN > 1 ? 1 : 0] = i1;
  b.out = out;
  b.s = s;
  b.aliased = (NIN > 0 && i0 == out) || (NIN > 1 && i1 == out);
  VecSplit<double, VBYTES> sp(out, n);
  b.head = sp.head; b.nvec = sp.nvec; b.tail = sp.tail;
  return RangeForLaunch<StreamBody<NIN, Op, VBYTES>, BLOCK, UNROLL>::run(I, b, b.nvec, bps);
}

template <int NIN, class Op>
int stream_entry(b200_instance* I, const char* where, const double* i0, const double* i1, double* out, double s, int64_t n) {
  B200_CHECK_INST(I, where);
  if (n < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (n == 0) return 0;
  if (!out || (NIN > 0 && !i0) || (NIN > 1 && !i1)) return b200_set_error(B200_EINVAL, where, "NULL array");
  const uintptr_t ph = reinterpret_cast<uintptr_t>(out) % 32;
  const bool same_phase = (NIN < 1 || reinterpret_cast<uintptr_t>(i0) % 32 == ph) && (NIN < 2 || reinterpret_cast<uintptr_t>(i1) % 32 == ph);
  const int bps = b200_tune("stream.bps", 0);
  if (!same_phase || ph % 8) return launch<NIN, Op, 8, 256, 4>(I, i0, i1, out, s, n, bps);
  const int vb = b200_tune("stream.vbytes", 32), un = b200_tune("stream.unroll", 2), bl = b200_tune("stream.block", 256);
#define CFG(V, B, U) if (vb == V && bl == B && un == U) return launch<NIN, Op, V, B, U>(I, i0, i1, out, s, n, bps);
  CFG(32, 256, 2)
#ifdef B200_SWEEP
  CFG(32, 256, 1) CFG(32, 256, 4) CFG(32, 512, 1) CFG(32, 512, 2) CFG(32, 128, 2) CFG(32, 128, 4)
  CFG(16, 256, 2) CFG(16, 256, 4) CFG(16, 512, 2) CFG(8, 256, 4) CFG(8, 256, 8) CFG(32, 1024, 1) CFG(16, 1024, 2)
#endif
#undef CFG
  return b200_set_error(B200_EUNSUPPORTED, where, "tuning combination not compiled in");
}
}  // namespace

extern "C" {
int b200_stream_set_f64(b200_instance* I, double* a, double value, int64_t n) { return stream_entry<0, SetOp>(I, "b200_stream_set_f64", nullptr, nullptr, a, value, n); }
int b200_stream_copy_f64(b200_instance* I, const double* a, double* b, int64_t n) { return stream_entry<1, CopyOp>(I, "b200_stream_copy_f64", a, nullptr, b, 0.0, n); }
int b200_stream_scale_f64(b200_instance* I, double* b, const double* c, double s, int64_t n) { return stream_entry<1, ScaleOp>(I, "b200_stream_scale_f64", c, nullptr, b, s, n); }
int b200_stream_add_f64(b200_instance* I, const double* a, const double* b, double* c, int64_t n) { return stream_entry<2, AddOp>(I, "b200_stream_add_f64", a, b, c, 0.0, n); }
int b200_stream_triad_f64(b200_instance* I, double* a, const double* b, const double* c, double s, int64_t n) { return stream_entry<2, TriadOp>(I, "b200_stream_triad_f64", b, c, a, s, n); }
}
