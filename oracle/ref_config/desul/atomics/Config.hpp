// hand-written stand-in for tpls/desul/Config.hpp.cmake.in (host-only build)
#ifndef DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_CONFIG_HPP_
#endif
