/*
 * pfhe_oracle.c -- CPU oracle for the primus-fhe polynomial-ring hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The shipped CUDA path (primus_fhe_b200/) never links or calls it.
 *
 * It is a plain-C restatement of the reference's *scalar* algorithms
 * (the reference is 100 % Rust and no Rust toolchain exists in this image, so
 * the reference itself cannot be compiled -- see DESIGN.md "Oracle").
 * Each function cites the reference file:line it follows, relative to
 * /root/reference/crates/.
 *
 * PARITY PINNING: the reference ships no golden vectors / KATs for this path
 * (every test is property-based on unseeded RNG or cross-implementation,
 * SURVEY.md 0.6 and 8c).  The oracle is therefore pinned by
 *   (1) the reference's own deterministic test inputs in primus_rns/tests/rns.rs,
 *   (2) the same cross-implementation / property checks the reference's tests
 *       make (Harvey table == generic table, round trips, lazy ranges,
 *       Shoup == Barrett, decomposition error bound and exact identity),
 *   (3) an independent pure-Python big-integer model (oracle/pymodel.py) and
 *       direct O(n^2) evaluation / schoolbook negacyclic products.
 * The NTT external product (no reference test at all) and blind rotation (not
 * in the reference) are "parity unpinned" at the composed level: no reference
 * output exists to compare with.  What stands in for it (tests/test_oracle.py):
 * the schoolbook identity at the BASELINE degrees, and functional tests with
 * real noisy LWE / RLWE / RGSW encryptions -- RLWE(m) x RGSW(mu) decrypts to
 * m * mu, and a programmable bootstrap returns LUT[m] -- at the config-4 / 5
 * parameters (tests/extprod_common.py, tests/bootstrap_common.py).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "pfhe_oracle.h"

typedef unsigned __int128 u128;
typedef __int128 s128;

#define OW 32
#define uw uint32_t
#define u2w uint64_t
#define s2w int64_t
#define ON(name) name##32
#include "oracle_impl.inc"
#undef OW
#undef uw
#undef u2w
#undef s2w
#undef ON

#define OW 64
#define uw uint64_t
#define u2w u128
#define s2w s128
#define ON(name) name##64
#include "oracle_impl.inc"
#undef OW
#undef uw
#undef u2w
#undef s2w
#undef ON

#ifdef _OPENMP
#include <omp.h>
int o_max_threads(void) { return omp_get_max_threads(); }
#else
int o_max_threads(void) { return 1; }
#endif
