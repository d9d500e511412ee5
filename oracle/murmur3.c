/* Oracle (TEST INFRASTRUCTURE, not product code): C restatement of the string
 * hash behind /root/reference/torchctr/utils.py:103-119
 *     hash_bucket(v, buckets, seed) = murmurhash3_32(str(v), seed, positive=True) % buckets
 * murmurhash3_32 is scikit-learn's wrapper of MurmurHash3_x86_32 (third-party,
 * scikit-learn>=1.5.1, not vendored in the reference); the published algorithm
 * is restated below.  Built by oracle/Makefile into oracle/_build/liboracle_ctr.so
 * and checked against tests/golden/hash_golden.json (sklearn 1.9.0 outputs).
 */
#include <stdint.h>
#include <stddef.h>

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t ctr_oracle_murmur3_32(const uint8_t *key, int len, uint32_t seed)
{
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = seed;
    int nblocks = len / 4;
    for (int i = 0; i < nblocks; ++i) {
        const uint8_t *p = key + 4 * i;
        uint32_t k = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        k *= c1; k = rotl32(k, 15); k *= c2;
        h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
    }
    const uint8_t *tail = key + 4 * nblocks;
    uint32_t k = 0;
    switch (len & 3) {
    case 3: k ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1: k ^= (uint32_t)tail[0];
            k *= c1; k = rotl32(k, 15); k *= c2; h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

/* decimal ASCII of a signed 64-bit id: what transformer.py:371-381 feeds the hash
 * for numeric categories.  Returns the length. */
int ctr_oracle_itoa(int64_t v, uint8_t out[20])
{
    uint8_t tmp[20];
    int n = 0, len = 0;
    uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    do { tmp[n++] = (uint8_t)('0' + u % 10u); u /= 10u; } while (u);
    if (v < 0) out[len++] = '-';
    while (n) out[len++] = tmp[--n];
    return len;
}

void ctr_oracle_hash_bucket_i64(const int64_t *ids, int64_t n, uint32_t buckets, uint32_t seed, int32_t *out)
{
    for (int64_t i = 0; i < n; ++i) {
        uint8_t buf[20];
        int len = ctr_oracle_itoa(ids[i], buf);
        out[i] = (int32_t)(ctr_oracle_murmur3_32(buf, len, seed) % buckets);
    }
}

/* Masked gather + sum pool of models/dnn.py:53-59 on a padded [B, L] id matrix
 * (negative id = pad).  mode 0 = sum (the reference), 1 = mean over valid ids
 * (extension, parity unpinned by the reference). Used as the single-thread CPU
 * "port" baseline for the lookup kernel. */
void ctr_oracle_pool(const int64_t *ids, int64_t B, int64_t L, const float *table, int64_t D,
                     int mode, float *out, int64_t out_stride)
{
    for (int64_t b = 0; b < B; ++b) {
        float *o = out + b * out_stride;
        for (int64_t d = 0; d < D; ++d) o[d] = 0.f;
        int64_t cnt = 0;
        for (int64_t l = 0; l < L; ++l) {
            int64_t id = ids[b * L + l];
            if (id < 0) continue;
            const float *row = table + id * D;
            for (int64_t d = 0; d < D; ++d) o[d] += row[d];
            ++cnt;
        }
        if (mode == 1 && cnt > 1) {
            float inv = 1.0f / (float)cnt;
            for (int64_t d = 0; d < D; ++d) o[d] *= inv;
        }
    }
}
