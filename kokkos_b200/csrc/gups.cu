// gups.cu -- Kokkos::atomic_* hot loops (C ABI): benchmarks/gups update loop
// (benchmarks/gups/gups.cpp:83-97) and a scatter-add of doubles.
//
// The reference reaches the hardware through desul: Kokkos::atomic_add ->
// desul::atomic_add(..., MemoryOrderRelaxed, MemoryScopeDevice) -> inline PTX
// `red.add.relaxed.gpu.global.*` behind an __isGlobal branch
// (core/src/Kokkos_Atomics_Desul_Wrapper.hpp:86-145,
//  tpls/desul/include/desul/atomics/cuda/cuda_cc7_asm_atomic_op.inc_isglobal:5-106).
// Here the no-return forms are emitted directly (kb200/Atomic.hpp): SASS RED.E.ADD.64.STRONG.GPU /
// RED.E.XOR.64 -- fire-and-forget L2 atomics, several independent updates in flight per thread.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/Atomic.hpp>
#include <kb200/impl/ForKernel.hpp>

using namespace kb200;
using namespace kb200::Impl;

namespace {
struct AddI64 { KB200_DEVICE_FUNCTION static void apply(int64* p, int64 v) { red_add_g((unsigned long long*)p, (unsigned long long)v); } };
struct XorI64 { KB200_DEVICE_FUNCTION static void apply(int64* p, int64 v) { red_xor_g((unsigned long long*)p, (unsigned long long)v); } };

template <class Op>
struct GupsBody {
  using packet = int64;  // the index: loads are batched ahead of the atomics
  int64* table;
  const int64* indices;
  int64 datum;
  KB200_DEVICE_FUNCTION packet load(int64 u) const { return __ldg(indices + u); }
  KB200_DEVICE_FUNCTION void store(const packet& idx, int64) const { Op::apply(table + idx, datum); }
  KB200_FUNCTION int64 edge_count() const { return 0; }
  KB200_DEVICE_FUNCTION void edge(int64) const {}
};
struct ScatterAddBody {
  struct packet { int64 idx; double v; };
  double* table;
  const int64* indices;
  const double* values;
  KB200_DEVICE_FUNCTION packet load(int64 u) const { return packet{__ldg(indices + u), __ldg(values + u)}; }
  KB200_DEVICE_FUNCTION void store(const packet& p, int64) const { red_add_g(table + p.idx, p.v); }
  KB200_FUNCTION int64 edge_count() const { return 0; }
  KB200_DEVICE_FUNCTION void edge(int64) const {}
};

template <class Op>
int gups_entry(b200_instance* I, const char* where, int64_t* table, int64_t len, const int64_t* idx, int64_t m, int64_t datum) {
  B200_CHECK_INST(I, where);
  if (m < 0 || len < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (m == 0) return 0;
  if (!table || !idx) return b200_set_error(B200_EINVAL, where, "NULL array");
  GupsBody<Op> b{(int64*)table, (const int64*)idx, (int64)datum};
  return RangeForLaunch<GupsBody<Op>, 256, 8>::run(I, b, m, b200_tune("gups.bps", 0));
}
}  // namespace

extern "C" {
int b200_gups_add_i64(b200_instance* I, int64_t* t, int64_t len, const int64_t* idx, int64_t m, int64_t d) { return gups_entry<AddI64>(I, "b200_gups_add_i64", t, len, idx, m, d); }
int b200_gups_xor_i64(b200_instance* I, int64_t* t, int64_t len, const int64_t* idx, int64_t m, int64_t d) { return gups_entry<XorI64>(I, "b200_gups_xor_i64", t, len, idx, m, d); }
int b200_atomic_add_f64(b200_instance* I, double* t, int64_t len, const int64_t* idx, const double* v, int64_t m) {
  const char* where = "b200_atomic_add_f64";
  B200_CHECK_INST(I, where);
  if (m < 0 || len < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (m == 0) return 0;
  if (!t || !idx || !v) return b200_set_error(B200_EINVAL, where, "NULL array");
  ScatterAddBody b{t, (const int64*)idx, v};
  return RangeForLaunch<ScatterAddBody, 256, 4>::run(I, b, m, 0);
}
}
