// hand-written stand-in for tpls/desul/Config.hpp.cmake.in (CUDA build, no relocatable device code)
#ifndef DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_ENABLE_CUDA
#endif
