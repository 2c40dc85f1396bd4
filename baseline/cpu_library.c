/* Best-effort CPU LIBRARY baseline for header_range_1024 (baseline B) -- not the oracle, not the product.
 * OpenSSL (libcrypto 3): SHA-256 with the CPU's SHA extensions and Ed25519 verification through EVP, OpenMP over ranges.
 * Per range it hashes messages of the sizes the witness schedule holds (20 969 digests: 1 024 header leaves of 35 bytes,
 * 1 024 of 73, 200 validator leaves of 45, the rest 65-byte inner nodes and tuple leaves) and verifies 100 signatures over
 * 108-byte votes.  It yields digests and accept / reject only -- none of the EC intermediates, quotients or the
 * request-order layout the witness needs -- so it bounds what tuned CPU libraries could do with the same inputs.
 * Built by __graft_entry__.build() into baseline/_cpulib/ (git-ignored); timed by bench.py as `cpu_library_baseline`. */
#include <omp.h>
#include <openssl/evp.h>
#include <openssl/sha.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_DIGESTS 20969
#define N_SIGS 100
#define VOTE_LEN 108

static uint32_t msg_size(int i) { return i < 1024 ? 35 : (i < 2048 ? 73 : (i < 2248 ? 45 : 65)); }

/* returns seconds of wall time for n_ranges ranges on `threads` threads, < 0 on failure; *sink defeats dead-code removal */
double cpulib_header_range(int n_ranges, int threads, uint32_t *sink) {
    static uint8_t msgs[N_DIGESTS * 80], votes[N_SIGS][VOTE_LEN], pks[N_SIGS][32], sigs[N_SIGS][64];
    uint32_t x = 0x9E3779B9u;
    for (size_t i = 0; i < sizeof msgs; i++) { x = x * 1664525u + 1013904223u; msgs[i] = (uint8_t)(x >> 24); }
    for (int i = 0; i < N_SIGS; i++) {
        uint8_t sk[32];
        for (int k = 0; k < 32; k++) { x = x * 1664525u + 1013904223u; sk[k] = (uint8_t)(x >> 24); }
        for (int k = 0; k < VOTE_LEN; k++) { x = x * 1664525u + 1013904223u; votes[i][k] = (uint8_t)(x >> 24); }
        EVP_PKEY *key = EVP_PKEY_new_raw_private_key(EVP_PKEY_ED25519, NULL, sk, 32);
        EVP_MD_CTX *c = EVP_MD_CTX_new();
        size_t sl = 64, pl = 32;
        if (!key || !c || EVP_DigestSignInit(c, NULL, NULL, NULL, key) != 1 || EVP_DigestSign(c, sigs[i], &sl, votes[i], VOTE_LEN) != 1 ||
            EVP_PKEY_get_raw_public_key(key, pks[i], &pl) != 1)
            return -1.0;
        EVP_MD_CTX_free(c);
        EVP_PKEY_free(key);
    }
    int bad = 0;
    uint32_t acc = 0;
    double t0 = 0, dt = 0;
#pragma omp parallel num_threads(threads) reduction(+ : bad, acc)
    {
        /* per-thread key objects and one verification context, created outside the timed region (OpenSSL 3 takes global
         * locks when it creates keys and fetches algorithms: with them inside the loop 8 threads ran no faster than one) */
        EVP_PKEY *keys[N_SIGS];
        for (int i = 0; i < N_SIGS; i++) keys[i] = EVP_PKEY_new_raw_public_key(EVP_PKEY_ED25519, NULL, pks[i], 32);
        EVP_MD_CTX *c = EVP_MD_CTX_new();
#pragma omp barrier
#pragma omp master
        t0 = omp_get_wtime();
#pragma omp for schedule(dynamic, 1)
        for (int r = 0; r < n_ranges; r++) {
            uint8_t d[32];
            for (int i = 0; i < N_DIGESTS; i++) {
                SHA256_CTX s;
                SHA256_Init(&s);
                SHA256_Update(&s, msgs + 80 * (size_t)i, msg_size(i));
                SHA256_Final(d, &s);
                acc += d[0];
            }
            for (int i = 0; i < N_SIGS; i++) {
                if (!keys[i] || !c || EVP_MD_CTX_reset(c) != 1 || EVP_DigestVerifyInit(c, NULL, NULL, NULL, keys[i]) != 1 ||
                    EVP_DigestVerify(c, sigs[i], 64, votes[i], VOTE_LEN) != 1)
                    bad++;
            }
        }
#pragma omp master
        dt = omp_get_wtime() - t0;
        EVP_MD_CTX_free(c);
        for (int i = 0; i < N_SIGS; i++) EVP_PKEY_free(keys[i]);
    }
    if (sink) *sink = acc;
    return bad ? -1.0 : dt;
}
