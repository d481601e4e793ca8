// common.cuh -- shared definitions for the dropest_b200 device pipeline (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace dge
{

struct CudaError : std::runtime_error
{
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

// malformed caller input (-> DGE_ERR_INVALID) and exhausted device tables (-> DGE_ERR_CAPACITY)
struct InvalidInput : std::runtime_error
{
    explicit InvalidInput(const std::string &m) : std::runtime_error(m) {}
};
struct CapacityError : std::runtime_error
{
    explicit CapacityError(const std::string &m) : std::runtime_error(m) {}
};

#define DGE_CUDA(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            throw ::dge::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                   std::to_string(__LINE__) + ")");                                            \
    } while (0)

#define DGE_LAUNCH_CHECK() DGE_CUDA(cudaGetLastError())

constexpr uint64_t EMPTY64 = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr uint32_t NO_GENE = 0xFFFFFFu;
constexpr uint32_t FLAG_UMI_N = 1u << 27, FLAG_CB_N = 1u << 28; // dge_record16.gene flags (include/dropest_b200.h)
constexpr uint64_t CB_N_BIT = 1ull << 40;                         // marks a barcode-table entry that is an index into the N-barcode list

// count | mark << 29  (count < 2^29 reads per UMI)
constexpr int VAL_MARK_SHIFT = 29;
constexpr uint32_t VAL_COUNT_MASK = (1u << VAL_MARK_SHIFT) - 1;

// One slot of the barcode table (open addressing, linear probing).  The slot index is the internal cell id.
struct __align__(16) CellSlot
{
    unsigned long long cb; // 2-bit packed barcode, EMPTY64 when free
    uint32_t first_idx;    // min read_idx over all reads of the barcode (cell id order of the reference, CellsDataContainer.cpp:64-69)
    uint32_t n_intergenic; // reads without a gene (CellsDataContainer.cpp:73-78)
};

// Layout of the 64-bit grouping key:  [ slot : tb | gene : gb | umi : ub | mark : 3 ]   (right aligned, kb = tb+gb+ub+3 <= 64)
// ukey = key >> 3 identifies one (cell, gene, UMI); sorting by ukey orders by cell slot, then gene id, then UMI value.
struct KeyLayout
{
    int tb, gb, ub, kb;
    int cbb; // 2 * cb_len: barcode bits of dge_record16.key >> 24 (record validation)
    int ul;  // 2 * umi_len: bits of a plain (N-free) UMI; with allow_n the UMI field is ub = max(ul, 20) + 1 bits wide and an N-UMI is stored as
             // [1 : index into the caller's N-UMI list], a barcode with N as [1 << 40 | index] in the barcode table
    int ne;  // allow_n
    __host__ __device__ uint64_t compose(uint32_t slot, uint32_t gene, uint32_t umi, uint32_t mark) const
    {
        return (((uint64_t(slot) << gb | gene) << ub | umi) << 3) | mark;
    }
    __host__ __device__ uint32_t slot_of_ukey(uint64_t ukey) const { return uint32_t(ukey >> (gb + ub)); }
    __host__ __device__ uint32_t gene_of_ukey(uint64_t ukey) const { return uint32_t(ukey >> ub) & ((1u << gb) - 1); }
    __host__ __device__ uint32_t umi_of_ukey(uint64_t ukey) const { return uint32_t(ukey & ((1ull << ub) - 1)); }
    __host__ __device__ uint64_t cg_of_ukey(uint64_t ukey) const { return ukey >> ub; }
    __host__ __device__ uint64_t gu_of_ukey(uint64_t ukey) const { return ukey & ((1ull << (gb + ub)) - 1); }
};

// A run of packed keys produced by one block of the fill kernel (k_fill_pipe): the L1 partition pass consumes a list of them.
struct KeyRegion
{
    const uint64_t *keys;
    uint32_t count;
    uint32_t tile0; // index of the region's first tile in the partition pass (k_region_tiles)
};

// One tile of the L1 partition pass over the key regions (k_region_tiles -> k_l1_scatter_regions)
struct KeyTile
{
    const uint64_t *keys;
    uint32_t count;
    uint32_t pad;
};

__host__ __device__ inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// Hash used for the barcode table AND for routing barcodes to ranks (host mirror: dropest_b200/synth.py:barcode_hash).
__host__ __device__ inline uint64_t barcode_hash(uint64_t cb) { return mix64(cb + 0x9E3779B97F4A7C15ull); }

inline int ceil_log2_u64(uint64_t x)
{
    int b = 0;
    while ((1ull << b) < x && b < 63) ++b;
    return b;
}

template <class T> inline T div_up(T a, T b) { return (a + b - 1) / b; }

// A grow-only device buffer.
struct DevBuf
{
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
    }
    void reserve(size_t n)
    {
        if (n <= bytes) return;
        release();
        size_t want = n + (n >> 4) + 256;
        DGE_CUDA(cudaMalloc(&p, want));
        bytes = want;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// A grow-only page-locked host buffer (D2H / H2D staging at full PCIe speed, truly asynchronous copies).
struct PinnedBuf
{
    void *p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    void reserve(size_t n)
    {
        if (n <= bytes) return;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        size_t want = n + (n >> 3) + 4096;
        DGE_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        bytes = want;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

} // namespace dge
