// TEST INFRASTRUCTURE ONLY — force-included (g++ -include) in front of every reference
// translation unit when building oracle/_ref/.
//
// The reference sources are compiled UNMODIFIED from /root/reference/src.  The only
// deviation from "as shipped" is made here, without touching them: we pre-define the
// include guard of src/base/constants.h (constants.h:1-2) and supply the same five
// constants ourselves, with `multithreaded = false`.
//
// Why: as shipped, `multithreaded = true` (constants.h:11) fans agents out over
// std::threads (economy.cpp:62-93) and the offer review/accept protocol has a
// check-then-act race (agent.cpp:118-161) that makes results non-deterministic.  The
// reference itself contains the single-threaded branch (economy.cpp:117-124); this
// selects it.  All other values are identical to constants.h:8-12.
#ifndef CONSTANTS_H
#define CONSTANTS_H
#include <thread>
namespace constants {
    const unsigned int verbose = 1;
    const double eps = 1e-8;
    const double largeNumber = 1e8;
#ifdef FASTACE_REF_AS_SHIPPED_MT
    const bool multithreaded = true;
#else
    const bool multithreaded = false;
#endif
    const unsigned int numThreads = std::thread::hardware_concurrency();
}
#endif
