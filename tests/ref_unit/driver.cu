// driver.cu -- main() for the reference's unit tests on the B200 execution space (the test headers are force-included by
// the Makefile from /root/reference/core/unit_test; see tests.list).  Mirrors core/unit_test/UnitTestMainInit.cpp.
#include <gtest/gtest.h>
#include <Kokkos_Core.hpp>

int main(int argc, char* argv[]) {
  Kokkos::initialize(argc, argv);
  ::testing::InitGoogleTest(&argc, argv);
  const int result = RUN_ALL_TESTS();
  Kokkos::finalize();
  return result;
}
