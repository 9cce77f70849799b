/*
 * oracle/gen.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Deterministic synthetic-table generator that writes tables in the reference's on-disk format
 * so the CPU oracle and the CUDA library scan identical files:
 *   <table>/meta.bin   write_table_meta   /root/reference/src/io/table_io.jl:9-19
 *   <table>/<id>.bin   make_column_file   /root/reference/src/io/filesystem.jl:14-23
 *   block framing      commit_block_write! /root/reference/src/io/BlockStreams.jl:36-60
 *                      (Int32 rows | Int64 origin | Int64 compressed | one raw LZ4 block, accel = 2)
 *   block bodies       write_block_body   /root/reference/src/io/blocks.jl:2-33
 * The compressor is the system liblz4 (dlopen "liblz4.so.1", LZ4_compress_fast -- the library
 * the reference binds through CodecLz4) when present, else oracle/lz4_ref.c.
 *
 * Values are a pure function of (seed, column index, block number), so any thread count and any
 * block range produce the same bytes.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <errno.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

int orc_lz4_compress_bound(int n);
int orc_lz4_compress(const uint8_t *src, uint8_t *dst, int n, int cap, int accel);

typedef int (*lz4_fast_fn)(const char *, char *, int, int, int);
static lz4_fast_fn g_lz4_fast;
static int g_lz4_probed;

static void probe_liblz4(void)
{
    if (g_lz4_probed) return;
    g_lz4_probed = 1;
    void *h = dlopen("liblz4.so.1", RTLD_NOW | RTLD_LOCAL);
    if (h) g_lz4_fast = (lz4_fast_fn)dlsym(h, "LZ4_compress_fast");
}

ORC_API int orc_have_liblz4(void) { probe_liblz4(); return g_lz4_fast != NULL; }

/* compress with the reference's codec call: LZ4_compress_fast(src, dst, n, bound, COMPRESSION_LEVEL=2) */
ORC_API int orc_compress_block(const uint8_t *src, uint8_t *dst, int n, int cap, int prefer_system)
{
    probe_liblz4();
    if (prefer_system && g_lz4_fast) return g_lz4_fast((const char *)src, (char *)dst, n, cap, 2);
    return orc_lz4_compress(src, dst, n, cap, 2);
}

/* ---- RNG: splitmix64 seeding + xoshiro256** ------------------------------------------------ */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed, uint64_t col, uint64_t block)
{
    uint64_t x = seed ^ (col * 0xD1342543DE82EF95ull) ^ (block * 0xA24BAED4963EE407ull);
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t *s = r->s, result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return (uint64_t)(((__uint128_t)rng_next(r) * n) >> 64); }
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---- column spec --------------------------------------------------------------------------- */
enum { D_IUNIFORM = 1, D_FUNIFORM, D_FGRID, D_ISEQ, D_BRANDS, D_DECIMAL };

typedef struct {
    char name[64];
    char typestr[64];
    int dist;
    double p[3];
    double missing_prob;
    int is_string, nullable, elsize, is_float;
} colspec;

static const char *k_brands[8] = {"apple", "samsung", "huawai", "microsoft", "dell", "xbox", "sony", "intel"};  /* docs/src/index.md:58 */

static int parse_spec(const char *spec, colspec *cols, int maxcols)
{
    int n = 0;
    char *dup = strdup(spec), *save1 = NULL;
    for (char *tok = strtok_r(dup, ";", &save1); tok; tok = strtok_r(NULL, ";", &save1)) {
        if (n >= maxcols) { free(dup); return -1; }
        colspec *c = &cols[n];
        memset(c, 0, sizeof *c);
        char *save2 = NULL;
        char *f = strtok_r(tok, ":", &save2);
        if (!f) { free(dup); return -1; }
        snprintf(c->name, sizeof c->name, "%s", f);
        f = strtok_r(NULL, ":", &save2);
        if (!f) { free(dup); return -1; }
        snprintf(c->typestr, sizeof c->typestr, "%s", f);
        f = strtok_r(NULL, ":", &save2);
        if (!f) { free(dup); return -1; }
        if (!strcmp(f, "iuniform")) c->dist = D_IUNIFORM;
        else if (!strcmp(f, "funiform")) c->dist = D_FUNIFORM;
        else if (!strcmp(f, "fgrid")) c->dist = D_FGRID;
        else if (!strcmp(f, "iseq")) c->dist = D_ISEQ;
        else if (!strcmp(f, "brands")) c->dist = D_BRANDS;
        else if (!strcmp(f, "decimal")) c->dist = D_DECIMAL;
        else { free(dup); return -1; }
        int k = 0;
        while ((f = strtok_r(NULL, ":", &save2))) {
            if (f[0] == 'm' && f[1] == '=') c->missing_prob = atof(f + 2);
            else if (k < 3) c->p[k++] = atof(f);
        }
        c->nullable = strncmp(c->typestr, "Missing(", 8) == 0;
        const char *base = c->nullable ? c->typestr + 8 : c->typestr;
        c->is_string = strncmp(base, "String", 6) == 0;
        c->is_float = strncmp(base, "Float64", 7) == 0;
        if (!c->is_string && strncmp(base, "Int64", 5) != 0 && !c->is_float) { free(dup); return -1; }
        c->elsize = c->is_string ? 0 : 8;
        n++;
    }
    free(dup);
    return n;
}


typedef struct {
    const colspec *col; uint64_t seed; int colidx; int64_t b0; int cnt; int64_t nrows, block_size; int bound, prefer_system;
    uint8_t **bodies, **comps; int64_t *origins, *csizes;
    int next;
} wave_job;

static int64_t make_body(const colspec *c, uint64_t seed, int colidx, int64_t block, int64_t row0, int64_t rows, uint8_t *body);

static void *wave_worker(void *arg)
{
    wave_job *j = arg;
    for (;;) {
        int k = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (k >= j->cnt) break;
        int64_t blk = j->b0 + k, row0 = blk * j->block_size;
        int64_t rows = j->nrows - row0 < j->block_size ? j->nrows - row0 : j->block_size;
        j->origins[k] = make_body(j->col, j->seed, j->colidx, blk, row0, rows, j->bodies[k]);
        j->csizes[k] = orc_compress_block(j->bodies[k], j->comps[k], (int)j->origins[k], j->bound, j->prefer_system);
    }
    return NULL;
}

/* body of one block, layouts per blocks.jl:2-33; returns body size */
static int64_t make_body(const colspec *c, uint64_t seed, int colidx, int64_t block, int64_t row0, int64_t rows, uint8_t *body)
{
    rng_t r;
    rng_seed(&r, seed, (uint64_t)colidx, (uint64_t)block);
    if (c->is_string) {
        int32_t *sizes = (int32_t *)(body + 4);
        uint8_t *chars = body + 4 + 4 * rows;
        int64_t pos = 0;
        for (int64_t i = 0; i < rows; i++) {
            int miss = c->nullable && rng_unit(&r) < c->missing_prob;
            char tmp[32];
            const char *s;
            int len;
            if (c->dist == D_BRANDS) { s = k_brands[rng_below(&r, 8)]; len = (int)strlen(s); }
            else { len = snprintf(tmp, sizeof tmp, "%d", (int32_t)(uint32_t)rng_next(&r)); s = tmp; }
            if (miss) { sizes[i] = -1; continue; }
            sizes[i] = len;
            memcpy(chars + pos, s, (size_t)len);
            pos += len;
        }
        int32_t ds = (int32_t)pos;
        memcpy(body, &ds, 4);
        return 4 + 4 * rows + pos;
    }
    int64_t nw = c->nullable ? (rows + 63) / 64 : 0;
    uint64_t *bits = (uint64_t *)body;
    uint8_t *vals = body + nw * 8;
    if (c->nullable) memset(bits, 0, (size_t)nw * 8);
    for (int64_t i = 0; i < rows; i++) {
        int miss = c->nullable && rng_unit(&r) < c->missing_prob;
        if (miss) bits[i >> 6] |= 1ull << (i & 63);
        if (c->is_float) {
            double v;
            if (c->dist == D_FGRID) {
                /* rand(lo:step:hi) */
                uint64_t steps = (uint64_t)((c->p[2] - c->p[0]) / c->p[1] + 0.5) + 1;
                v = c->p[0] + (double)rng_below(&r, steps) * c->p[1];
            } else v = rng_unit(&r);
            if (miss) { uint64_t g = 0xDEADBEEFCAFEF00Dull ^ (uint64_t)i; memcpy(vals + 8 * i, &g, 8); }   /* garbage under missing (missings.jl:1) */
            else memcpy(vals + 8 * i, &v, 8);
        } else {
            int64_t v;
            if (c->dist == D_ISEQ) v = row0 + i + 1;
            else v = (int64_t)c->p[0] + (int64_t)rng_below(&r, (uint64_t)((int64_t)c->p[1] - (int64_t)c->p[0] + 1));
            if (miss) v = (int64_t)(0x5A5A5A5A00000000ull | (uint64_t)i);
            memcpy(vals + 8 * i, &v, 8);
        }
    }
    return nw * 8 + rows * 8;
}

static void put_string(FILE *f, const char *s)
{
    int32_t n = (int32_t)strlen(s);
    fwrite(&n, 4, 1, f);
    fwrite(s, 1, (size_t)n, f);
}

/*
 * spec: "name:Type:dist[:p0[:p1[:p2]]][:m=prob];..." ; dist in {iuniform lo hi, funiform, fgrid lo step hi,
 * iseq, brands, decimal}.  Returns 0 on success.  stats[0] = uncompressed bytes, stats[1] = compressed bytes.
 */
ORC_API int orc_gen_table(const char *path, const char *spec, int64_t nrows, int64_t block_size, uint64_t seed,
                          int nthreads, int prefer_system_lz4, int64_t *stats)
{
    colspec cols[64];
    int ncols = parse_spec(spec, cols, 64);
    if (ncols <= 0 || block_size <= 0 || nrows < 0) return -1;
    if (mkdir(path, 0777) && errno != EEXIST) return -2;
    char p[1200];
    snprintf(p, sizeof p, "%s/meta.bin", path);
    FILE *mf = fopen(p, "wb");
    if (!mf) return -3;
    int64_t ver = 1, nc = ncols;
    fwrite(&ver, 8, 1, mf); fwrite(&block_size, 8, 1, mf); fwrite(&nc, 8, 1, mf);
    for (int i = 0; i < ncols; i++) {
        int64_t id = i + 1;
        fwrite(&id, 8, 1, mf);
        put_string(mf, cols[i].name);
        put_string(mf, cols[i].typestr);
    }
    fclose(mf);

    int64_t nblocks = (nrows + block_size - 1) / block_size;
    if (nthreads < 1) nthreads = 1;
    int wave = nthreads * 4;
    int64_t max_body = 4 + block_size * 4 + block_size * 16 + 64;
    if (max_body < block_size * 8 + (block_size / 64 + 1) * 8 + 64) max_body = block_size * 8 + (block_size / 64 + 1) * 8 + 64;
    int bound = orc_lz4_compress_bound((int)max_body);
    uint8_t **bodies = calloc((size_t)wave, sizeof *bodies), **comps = calloc((size_t)wave, sizeof *comps);
    int64_t *origins = calloc((size_t)wave, 8), *csizes = calloc((size_t)wave, 8);
    for (int i = 0; i < wave; i++) { bodies[i] = malloc((size_t)max_body); comps[i] = malloc((size_t)bound); }
    int64_t tot_u = 0, tot_c = 0;
    int rc = 0;
    for (int ci = 0; ci < ncols && !rc; ci++) {
        snprintf(p, sizeof p, "%s/%d.bin", path, ci + 1);
        FILE *f = fopen(p, "wb");
        if (!f) { rc = -4; break; }
        fwrite(&block_size, 8, 1, f);
        put_string(f, cols[ci].typestr);
        for (int64_t b0 = 0; b0 < nblocks; b0 += wave) {
            int cnt = (int)((nblocks - b0) < wave ? (nblocks - b0) : wave);
            wave_job job = { &cols[ci], seed, ci, b0, cnt, nrows, block_size, bound, prefer_system_lz4, bodies, comps, origins, csizes, 0 };
            int nt = nthreads < cnt ? nthreads : cnt;
            pthread_t th[256];
            if (nt > 256) nt = 256;
            for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, wave_worker, &job);
            wave_worker(&job);
            for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
            for (int k = 0; k < cnt; k++) {
                int64_t blk = b0 + k, row0 = blk * block_size;
                int32_t rows = (int32_t)(nrows - row0 < block_size ? nrows - row0 : block_size);
                if (csizes[k] <= 0) { rc = -5; break; }
                fwrite(&rows, 4, 1, f); fwrite(&origins[k], 8, 1, f); fwrite(&csizes[k], 8, 1, f);
                if (fwrite(comps[k], 1, (size_t)csizes[k], f) != (size_t)csizes[k]) { rc = -6; break; }
                tot_u += origins[k]; tot_c += csizes[k];
            }
            if (rc) break;
        }
        if (fclose(f)) rc = rc ? rc : -7;
    }
    for (int i = 0; i < wave; i++) { free(bodies[i]); free(comps[i]); }
    free(bodies); free(comps); free(origins); free(csizes);
    if (stats) { stats[0] = tot_u; stats[1] = tot_c; }
    return rc;
}
