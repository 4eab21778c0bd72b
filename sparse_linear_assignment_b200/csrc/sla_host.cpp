// sla_host.cpp -- libsla_host.so: the host-only entry points of include/sla.h, built with g++ (no CUDA runtime), for
// processes that must not map the CUDA library (bench.py --impl reference, CPU-only test runs).
#include "sla_host_impl.h"
