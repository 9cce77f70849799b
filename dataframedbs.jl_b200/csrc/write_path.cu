// write_path.cu -- the write path of a column file on the device.
//
// Stands in for make_column_file (/root/reference/src/io/filesystem.jl:22-31), write_column (src/io/columns.jl:65-84) and
// commit_block_write! (src/io/BlockStreams.jl:36-60): the column is cut into blocks of block_size rows, every block's body
// is laid out the way src/io/blocks.jl:2-33 writes it, compressed as ONE raw LZ4 block (lz4_compress.cu) and framed as
// Int32 rows | Int64 origin | Int64 compressed | payload.  Body assembly, compression and the compaction of the payloads run
// on the device; the host frames and writes.  The reference's reader opens the result unchanged.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace dfdb;

namespace {

constexpr int PACK_THREADS = 256;

struct PackArgs {
    int nblocks, elsize, nullable, is_string;
    int64_t block_size, nrows;
    const uint8_t *values;        // fixed width: nrows * elsize
    const uint8_t *missing;       // nrows bytes or null
    const int32_t *sizes;         // strings
    const uint8_t *chars;
    const int64_t *char_off;      // strings: first char of each block (nblocks + 1)
    uint8_t *bodies;
    const int64_t *body_off;
};

// one CTA per block: body layout of src/io/blocks.jl:2-33
__global__ void __launch_bounds__(PACK_THREADS) pack_bodies_kernel(const PackArgs A)
{
    for (int b = blockIdx.x; b < A.nblocks; b += gridDim.x) {
        const int64_t r0 = (int64_t)b * A.block_size;
        const int64_t rows = A.nrows - r0 < A.block_size ? A.nrows - r0 : A.block_size;
        uint8_t *body = A.bodies + A.body_off[b];
        if (A.is_string) {
            const int64_t c0 = A.char_off[b], c1 = A.char_off[b + 1];
            int32_t *head = reinterpret_cast<int32_t *>(body);
            if (threadIdx.x == 0) head[0] = (int32_t)(c1 - c0);
            for (int64_t r = threadIdx.x; r < rows; r += PACK_THREADS) head[1 + r] = A.sizes[r0 + r];
            uint8_t *dst = body + 4 + 4 * rows;
            for (int64_t i = threadIdx.x; i < c1 - c0; i += PACK_THREADS) dst[i] = A.chars[c0 + i];
        } else {
            uint8_t *vals = body;
            if (A.nullable) {
                // BitArray chunks: bit (i & 63) of word (i >> 6) set = missing; pad bits zero (read!(io, ::BitArray) checks them)
                const int64_t words = (rows + 63) >> 6;
                unsigned long long *bm = reinterpret_cast<unsigned long long *>(body);
                for (int64_t w = threadIdx.x; w < words; w += PACK_THREADS) {
                    unsigned long long x = 0;
                    for (int k = 0; k < 64; k++) {
                        const int64_t r = w * 64 + k;
                        if (r < rows && A.missing[r0 + r]) x |= 1ull << k;
                    }
                    bm[w] = x;
                }
                vals = body + words * 8;
            }
            const int64_t nbytes = rows * A.elsize;
            const uint8_t *src = A.values + r0 * A.elsize;
            for (int64_t i = threadIdx.x; i < nbytes; i += PACK_THREADS) vals[i] = src[i];
        }
    }
}

// payloads back to back: one CTA per block
__global__ void __launch_bounds__(PACK_THREADS) compact_payloads_kernel(const uint8_t *slots, const int64_t *slot_off, const int64_t *len, const int64_t *dst_off,
                                                                         uint8_t *dst, int nblocks)
{
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const uint8_t *s = slots + slot_off[b];
        uint8_t *d = dst + dst_off[b];
        for (int64_t i = threadIdx.x; i < len[b]; i += PACK_THREADS) d[i] = s[i];
    }
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    template <typename T> T *as() { return static_cast<T *>(p); }
    bool alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 16)) == cudaSuccess; }
};

int64_t compress_bound(int64_t n) { return n + n / 255 + 16; }

#define W_TRY(expr)                                                                                            \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) return fail(DFDB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));     \
    } while (0)

// compress n bodies that are resident on the device; payloads come back compacted in host vector `packed`
int compress_device_bodies(const uint8_t *d_bodies, const std::vector<int64_t> &body_off, const std::vector<int64_t> &body_len,
                           std::vector<int64_t> &comp_len, std::vector<uint8_t> &packed, std::vector<int64_t> &packed_off)
{
    const RuntimeView rv = runtime_view();
    cudaStream_t stream = static_cast<cudaStream_t>(rv.stream);
    const int n = (int)body_len.size();
    comp_len.assign((size_t)n, 0);
    packed_off.assign((size_t)n + 1, 0);
    if (n == 0) return DFDB_OK;
    std::vector<int64_t> slot_off((size_t)n);
    int64_t spos = 0;
    for (int i = 0; i < n; i++) { slot_off[(size_t)i] = spos; spos += (compress_bound(body_len[(size_t)i]) + 15) & ~(int64_t)15; }
    DevBuf d_slots, d_boff, d_blen, d_soff, d_clen, d_poff, d_packed;
    if (!d_slots.alloc((size_t)spos) || !d_boff.alloc((size_t)n * 8) || !d_blen.alloc((size_t)n * 8) || !d_soff.alloc((size_t)n * 8) ||
        !d_clen.alloc((size_t)n * 8) || !d_poff.alloc((size_t)n * 8))
        return fail(DFDB_ERR_NOMEM, "out of device memory for the compressor");
    W_TRY(cudaMemcpyAsync(d_boff.p, body_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    W_TRY(cudaMemcpyAsync(d_blen.p, body_len.data(), (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    W_TRY(cudaMemcpyAsync(d_soff.p, slot_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    CompressArgs ca{d_bodies, d_boff.as<int64_t>(), d_blen.as<int64_t>(), d_slots.as<uint8_t>(), d_soff.as<int64_t>(), d_clen.as<int64_t>(), n};
    if (launch_lz4_compress(ca, rv.counter, rv.sm_count, stream) != 0) return fail(DFDB_ERR_CUDA, "compress launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    runtime_count_launch();
    W_TRY(cudaMemcpyAsync(comp_len.data(), d_clen.p, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
    W_TRY(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; i++) packed_off[(size_t)i + 1] = packed_off[(size_t)i] + comp_len[(size_t)i];
    const int64_t total = packed_off[(size_t)n];
    if (!d_packed.alloc((size_t)total)) return fail(DFDB_ERR_NOMEM, "out of device memory for the compressed payloads");
    W_TRY(cudaMemcpyAsync(d_poff.p, packed_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    compact_payloads_kernel<<<std::min(n, rv.sm_count * 8), PACK_THREADS, 0, stream>>>(d_slots.as<uint8_t>(), d_soff.as<int64_t>(), d_clen.as<int64_t>(),
                                                                                       d_poff.as<int64_t>(), d_packed.as<uint8_t>(), n);
    if (cudaGetLastError() != cudaSuccess) return fail(DFDB_ERR_CUDA, "compaction launch failed");
    runtime_count_launch();
    packed.resize((size_t)std::max<int64_t>(total, 1));
    W_TRY(cudaMemcpyAsync(packed.data(), d_packed.p, (size_t)total, cudaMemcpyDeviceToHost, stream));
    W_TRY(cudaStreamSynchronize(stream));
    return DFDB_OK;
}

bool write_all(int fd, const void *p, size_t n)
{
    const char *c = static_cast<const char *>(p);
    while (n > 0) {
        ssize_t w = write(fd, c, n);
        if (w <= 0) return false;
        c += w;
        n -= (size_t)w;
    }
    return true;
}

}  // namespace

extern "C" {

int32_t dfdb_lz4_compress_blocks(const uint8_t *bodies, const int64_t *body_off, const int64_t *body_len, int32_t n, uint8_t *out,
                                 const int64_t *out_off, int64_t *comp_len)
{
    if (!runtime_view().inited) return fail(DFDB_ERR_CUDA, "dfdb_init has not been called (or no CUDA device is usable)");
    if (n <= 0) return DFDB_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(runtime_view().stream);
    std::vector<int64_t> boff((size_t)n), blen((size_t)n);
    int64_t pos = 0;
    for (int i = 0; i < n; i++) {
        if (body_len[i] < 0 || body_len[i] > 0x7E000000LL) return fail(DFDB_ERR_ARGUMENT, "bad body size");
        boff[(size_t)i] = pos; blen[(size_t)i] = body_len[i];
        pos += ((body_len[i] + 15) & ~(int64_t)15) + 16;
    }
    std::vector<uint8_t> staged((size_t)pos + 16, 0);
    for (int i = 0; i < n; i++) memcpy(staged.data() + boff[(size_t)i], bodies + body_off[i], (size_t)body_len[i]);
    DevBuf d_bodies;
    if (!d_bodies.alloc(staged.size())) return fail(DFDB_ERR_NOMEM, "out of device memory");
    W_TRY(cudaMemcpyAsync(d_bodies.p, staged.data(), staged.size(), cudaMemcpyHostToDevice, stream));
    std::vector<int64_t> clen, poff;
    std::vector<uint8_t> packed;
    int rc = compress_device_bodies(d_bodies.as<uint8_t>(), boff, blen, clen, packed, poff);
    if (rc) return rc;
    for (int i = 0; i < n; i++) {
        comp_len[i] = clen[(size_t)i];
        memcpy(out + out_off[i], packed.data() + poff[(size_t)i], (size_t)clen[(size_t)i]);
    }
    return DFDB_OK;
}

int32_t dfdb_write_table_meta(const char *table_path, int64_t block_size, int32_t ncols, const int64_t *ids, const char *const *names,
                              const char *const *typestrings)
{
    if (!table_path || block_size <= 0 || ncols < 0) return fail(DFDB_ERR_ARGUMENT, "bad table description");
    struct stat st;
    if (stat(table_path, &st) != 0 && mkdir(table_path, 0777) != 0) return fail(DFDB_ERR_IO, "cannot create %s", table_path);
    for (int i = 0; i < ncols; i++) {
        ColType ct;
        int rc = parse_typestring(typestrings[i], strlen(typestrings[i]), &ct);
        if (rc) return rc;
    }
    const std::string tmp = std::string(table_path) + "/meta.bin.tmp", dst = std::string(table_path) + "/meta.bin";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(DFDB_ERR_IO, "cannot write %s", tmp.c_str());
    const int64_t version = 1, nc = ncols;
    bool ok = fwrite(&version, 8, 1, f) == 1 && fwrite(&block_size, 8, 1, f) == 1 && fwrite(&nc, 8, 1, f) == 1;
    for (int i = 0; ok && i < ncols; i++) {
        const int32_t ln = (int32_t)strlen(names[i]), lt = (int32_t)strlen(typestrings[i]);
        ok = fwrite(&ids[i], 8, 1, f) == 1 && fwrite(&ln, 4, 1, f) == 1 && fwrite(names[i], 1, (size_t)ln, f) == (size_t)ln &&
             fwrite(&lt, 4, 1, f) == 1 && fwrite(typestrings[i], 1, (size_t)lt, f) == (size_t)lt;
    }
    ok = fclose(f) == 0 && ok;
    if (!ok || rename(tmp.c_str(), dst.c_str()) != 0) { unlink(tmp.c_str()); return fail(DFDB_ERR_IO, "cannot write %s", dst.c_str()); }
    return DFDB_OK;
}

int32_t dfdb_write_column_file(const char *table_path, int64_t col_id, const char *typestring, int64_t block_size, int64_t nrows, const void *values,
                               const uint8_t *missing, const int32_t *str_sizes, const uint8_t *str_chars, int64_t nchars, int64_t *compressed_total,
                               int64_t *uncompressed_total)
{
    if (!runtime_view().inited) return fail(DFDB_ERR_CUDA, "dfdb_init has not been called (or no CUDA device is usable)");
    if (!table_path || !typestring || block_size <= 0 || nrows < 0) return fail(DFDB_ERR_ARGUMENT, "bad column description");
    ColType ct;
    int rc = parse_typestring(typestring, strlen(typestring), &ct);
    if (rc) return rc;
    const bool is_string = ct.kind == DFDB_STRING;
    if (ct.kind == DFDB_TUPLE) return fail(DFDB_ERR_UNSUPPORTED, "tuple columns are not written by the device path");
    if (is_string ? (!str_sizes && nrows > 0) : (!values && nrows > 0)) return fail(DFDB_ERR_ARGUMENT, "missing input buffer");
    if (ct.nullable && !is_string && !missing && nrows > 0) return fail(DFDB_ERR_ARGUMENT, "nullable column without missing flags");
    const std::string path = std::string(table_path) + "/" + std::to_string(col_id) + ".bin";
    struct stat st;
    if (stat(path.c_str(), &st) == 0) return fail(DFDB_ERR_IO, "Column file with id %lld already exists", (long long)col_id);   // filesystem.jl:24
    cudaStream_t stream = static_cast<cudaStream_t>(runtime_view().stream);
    const int nblocks = (int)((nrows + block_size - 1) / block_size);
    // ---- block geometry on the host: body sizes, char ranges of String blocks ----
    std::vector<int64_t> body_off((size_t)nblocks), body_len((size_t)nblocks), char_off((size_t)nblocks + 1, 0);
    std::vector<int32_t> rows_b((size_t)nblocks);
    int64_t bpos = 0, csum = 0;
    for (int b = 0; b < nblocks; b++) {
        const int64_t r0 = (int64_t)b * block_size, rows = std::min(block_size, nrows - r0);
        rows_b[(size_t)b] = (int32_t)rows;
        int64_t len;
        if (is_string) {
            char_off[(size_t)b] = csum;
            for (int64_t r = r0; r < r0 + rows; r++) {
                const int32_t sz = str_sizes[r];
                if (sz < -1 || (sz == -1 && !ct.nullable)) return fail(DFDB_ERR_ARGUMENT, "bad string size %d at row %lld", sz, (long long)r);
                if (sz > 0) csum += sz;
            }
            len = 4 + 4 * rows + (csum - char_off[(size_t)b]);
        } else {
            len = rows * ct.elsize + (ct.nullable ? ((rows + 63) / 64) * 8 : 0);
        }
        if (len > 0x7E000000LL) return fail(DFDB_ERR_ARGUMENT, "block body of %lld bytes is too large for one LZ4 block", (long long)len);
        body_off[(size_t)b] = bpos;
        body_len[(size_t)b] = len;
        bpos += ((len + 15) & ~(int64_t)15) + 16;
    }
    char_off[(size_t)nblocks] = csum;
    if (is_string && csum != nchars) return fail(DFDB_ERR_ARGUMENT, "string sizes add up to %lld bytes, %lld given", (long long)csum, (long long)nchars);
    // ---- device: inputs -> bodies (src/io/blocks.jl layouts) -> LZ4 blocks -> compacted payloads ----
    std::vector<int64_t> comp_len, packed_off;
    std::vector<uint8_t> packed;
    if (nblocks > 0) {
        DevBuf d_vals, d_miss, d_sizes, d_chars, d_coff, d_bodies, d_boff;
        if (!d_bodies.alloc((size_t)bpos + 16) || !d_boff.alloc((size_t)nblocks * 8)) return fail(DFDB_ERR_NOMEM, "out of device memory for the block bodies");
        W_TRY(cudaMemsetAsync(d_bodies.p, 0, (size_t)bpos + 16, stream));
        W_TRY(cudaMemcpyAsync(d_boff.p, body_off.data(), (size_t)nblocks * 8, cudaMemcpyHostToDevice, stream));
        PackArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.nblocks = nblocks; pa.elsize = ct.elsize; pa.nullable = ct.nullable ? 1 : 0; pa.is_string = is_string ? 1 : 0;
        pa.block_size = block_size; pa.nrows = nrows;
        pa.bodies = d_bodies.as<uint8_t>(); pa.body_off = d_boff.as<int64_t>();
        if (is_string) {
            if (!d_sizes.alloc((size_t)nrows * 4) || !d_chars.alloc((size_t)nchars) || !d_coff.alloc((size_t)(nblocks + 1) * 8)) return fail(DFDB_ERR_NOMEM, "out of device memory");
            W_TRY(cudaMemcpyAsync(d_sizes.p, str_sizes, (size_t)nrows * 4, cudaMemcpyHostToDevice, stream));
            if (nchars > 0) W_TRY(cudaMemcpyAsync(d_chars.p, str_chars, (size_t)nchars, cudaMemcpyHostToDevice, stream));
            W_TRY(cudaMemcpyAsync(d_coff.p, char_off.data(), (size_t)(nblocks + 1) * 8, cudaMemcpyHostToDevice, stream));
            pa.sizes = d_sizes.as<int32_t>(); pa.chars = d_chars.as<uint8_t>(); pa.char_off = d_coff.as<int64_t>();
        } else {
            if (!d_vals.alloc((size_t)nrows * ct.elsize)) return fail(DFDB_ERR_NOMEM, "out of device memory");
            W_TRY(cudaMemcpyAsync(d_vals.p, values, (size_t)nrows * ct.elsize, cudaMemcpyHostToDevice, stream));
            pa.values = d_vals.as<uint8_t>();
            if (ct.nullable) {
                if (!d_miss.alloc((size_t)nrows)) return fail(DFDB_ERR_NOMEM, "out of device memory");
                W_TRY(cudaMemcpyAsync(d_miss.p, missing, (size_t)nrows, cudaMemcpyHostToDevice, stream));
                pa.missing = d_miss.as<uint8_t>();
            }
        }
        pack_bodies_kernel<<<std::min(nblocks, runtime_view().sm_count * 8), PACK_THREADS, 0, stream>>>(pa);
        if (cudaGetLastError() != cudaSuccess) return fail(DFDB_ERR_CUDA, "body assembly launch failed");
        runtime_count_launch();
        rc = compress_device_bodies(d_bodies.as<uint8_t>(), body_off, body_len, comp_len, packed, packed_off);
        if (rc) return rc;
    }
    // ---- host: frame and write (make_column_file header, then one frame per block; empty bodies are not written, BlockStreams.jl:38) ----
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0666);
    if (fd < 0) return fail(DFDB_ERR_IO, "cannot create %s", path.c_str());
    const int32_t tl = (int32_t)strlen(typestring);
    bool ok = write_all(fd, &block_size, 8) && write_all(fd, &tl, 4) && write_all(fd, typestring, (size_t)tl);
    int64_t ctot = 0, utot = 0;
    for (int b = 0; ok && b < nblocks; b++) {
        if (body_len[(size_t)b] == 0) continue;
        struct __attribute__((packed)) { int32_t rows; int64_t origin, compressed; } h{rows_b[(size_t)b], body_len[(size_t)b], comp_len[(size_t)b]};
        ok = write_all(fd, &h, 20) && write_all(fd, packed.data() + packed_off[(size_t)b], (size_t)comp_len[(size_t)b]);
        ctot += comp_len[(size_t)b];
        utot += body_len[(size_t)b];
    }
    ok = close(fd) == 0 && ok;
    if (!ok) { unlink(path.c_str()); return fail(DFDB_ERR_IO, "short write to %s", path.c_str()); }
    if (compressed_total) *compressed_total = ctot;
    if (uncompressed_total) *uncompressed_total = utot;
    return DFDB_OK;
}

}  // extern "C"
