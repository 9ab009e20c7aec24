/* c_abi_smoke.c -- the drop-in boundary used from plain C: no Python, no torch, only include/pfhe.h and libpfhe_cuda.so.
 *
 *   gcc -O2 -Iinclude examples/c_abi_smoke.c -o c_abi_smoke -Lprimus_fhe_b200/lib -lpfhe_cuda -Wl,-rpath,$PWD/primus_fhe_b200/lib
 *
 * Host slices in, host slices out (the shape of NttTable::transform_slice / inverse_transform_slice,
 * primus_ntt/src/ntt/mod.rs:16-113): forward + inverse must give the input back, and the fused product of a(X) with the monomial X must be
 * the negacyclic shift of a.  Exit code 0 and "c-abi smoke ok" on success. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pfhe.h"

#define CHECK(call)                                                                                           \
    do {                                                                                                      \
        pfhe_status s_ = (call);                                                                              \
        if (s_ != PFHE_OK) {                                                                                  \
            fprintf(stderr, "%s failed: %s %s\n", #call, pfhe_status_string(s_), pfhe_last_cuda_error());   \
            return 1;                                                                                         \
        }                                                                                                     \
    } while (0)

int main(void) {
    const uint64_t q = 1125899906826241ull; /* benches/bench_u64.rs:8 */
    const uint32_t log_n = 12;
    const size_t n = (size_t)1 << log_n, batch = 64;
    int devices = 0;
    CHECK(pfhe_device_count(&devices));
    if (devices < 1) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    pfhe_ntt64 *t = NULL;
    CHECK(pfhe_ntt64_create(0, log_n, q, &t));
    uint64_t *a = malloc(batch * n * sizeof(uint64_t)), *ref = malloc(batch * n * sizeof(uint64_t));
    uint64_t *x = calloc(batch * n, sizeof(uint64_t)), *c = malloc(batch * n * sizeof(uint64_t));
    uint64_t state = 88172645463325252ull;
    for (size_t i = 0; i < batch * n; i++) { /* xorshift64 */
        state ^= state << 13, state ^= state >> 7, state ^= state << 17;
        a[i] = state % q;
    }
    memcpy(ref, a, batch * n * sizeof(uint64_t));
    const uint64_t launches_before = pfhe_launch_count();
    CHECK(pfhe_ntt64_transform_slices(t, a, batch, 0));
    if (memcmp(a, ref, batch * n * sizeof(uint64_t)) == 0) {
        fprintf(stderr, "forward transform left the data unchanged\n");
        return 1;
    }
    CHECK(pfhe_ntt64_inverse_transform_slices(t, a, batch, 0));
    if (memcmp(a, ref, batch * n * sizeof(uint64_t)) != 0) {
        fprintf(stderr, "round trip mismatch\n");
        return 1;
    }
    for (size_t p = 0; p < batch; p++) x[p * n + 1] = 1; /* the monomial X */
    CHECK(pfhe_ntt64_polymul_slices(t, ref, x, c, batch));
    for (size_t p = 0; p < batch; p++)
        for (size_t i = 0; i < n; i++) {
            const uint64_t want = i == 0 ? (ref[p * n + n - 1] ? q - ref[p * n + n - 1] : 0) : ref[p * n + i - 1];
            if (c[p * n + i] != want) {
                fprintf(stderr, "product mismatch at poly %zu coeff %zu\n", p, i);
                return 1;
            }
        }
    const uint64_t launched = pfhe_launch_count() - launches_before;
    pfhe_ntt64_destroy(t);
    free(a), free(ref), free(x), free(c);
    if (launched == 0) {
        fprintf(stderr, "no kernel launch was counted\n");
        return 1;
    }
    printf("c-abi smoke ok: %llu kernel launches, %d device(s)\n", (unsigned long long)launched, devices);
    return 0;
}
