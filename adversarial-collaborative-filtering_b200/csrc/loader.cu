// N1: device-side data loader.  `*.rating` TSV text -> (uid, iid, rating) columns -> the structures the hot
// path consumes: the de-duplicated (u, i) pair list in file order (Dataset.py:278-304 `trainMatrix` keys ->
// APR.py:30-36 `sampling`), the reference's `trainList` row of every line incl. its cursor quirk (Dataset.py:306-325,
// SURVEY B.4) and the sorted, de-duplicated CSR of trainList (negative rejection APR.py:77, evaluation candidates
// utils.py:211).  He-format `.test.negative` lines ("(u,i)\tn1\t...\tn99", Dataset.py:161-172) -> CSR of negatives.
//
// The file is read once by the host (one read() into pinned memory), everything else happens on the device:
//   tsv_count_kernel   line starts per 1 KB block of text                         (streaming, HBM bound)
//   cub ExclusiveSum   block bases
//   tsv_parse_kernel   block-local line numbering (cub::BlockScan) + field parsing, one thread per line start
//   cursor: trainList row = min(k + 1, k + prefix_min(uid_j - j)) for a uid-sorted file (two scans instead of the
//           reference's sequential cursor; proof by exhaustive check in tests/test_host_logic.py)
//   sort / unique / histogram for the CSR and the first-occurrence pair list (cub radix sort + select)
// Sorting and scans are CUB library calls; this is load-time work, not the training hot path.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace apr {

constexpr int kTsvBlock = 1024;   // bytes of text per thread block (256 threads x 4 bytes)

__device__ __forceinline__ bool line_start_at(const char* __restrict__ t, int64_t b, int64_t n) {
  if (b >= n) return false;
  const char c = t[b];
  if (c == '\n' || c == '\r') return false;                 // empty line (the reference skips "" and "\n")
  return b == 0 || t[b - 1] == '\n';
}

__global__ void __launch_bounds__(256) tsv_count_kernel(const char* __restrict__ text, int64_t n, int64_t* __restrict__ counts) {
  const int64_t b0 = int64_t(blockIdx.x) * kTsvBlock + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) c += line_start_at(text, b0 + k, n) ? 1 : 0;
  typedef cub::BlockReduce<int, 256> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  const int total = Reduce(tmp).Sum(c);
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

// unsigned decimal integer at t[p...]; stops at the first non-digit; ok=false if no digit or > INT32_MAX
__device__ __forceinline__ int64_t parse_uint(const char* __restrict__ t, int64_t& p, int64_t n, bool& ok) {
  int64_t v = 0;
  int digits = 0;
  while (p < n && t[p] >= '0' && t[p] <= '9') {
    v = v * 10 + (t[p] - '0');
    if (v > 0x7fffffffLL) ok = false;
    ++p; ++digits;
  }
  if (!digits) ok = false;
  return v;
}

// decimal float: [+-]digits[.digits][(e|E)[+-]digits] -> double (exact for the 1-7 significant digits ratings carry)
__device__ __forceinline__ double parse_float(const char* __restrict__ t, int64_t& p, int64_t n, bool& ok) {
  bool neg = false;
  if (p < n && (t[p] == '-' || t[p] == '+')) { neg = t[p] == '-'; ++p; }
  double m = 0.0;
  int digits = 0, scale = 0;
  while (p < n && t[p] >= '0' && t[p] <= '9') { m = m * 10.0 + (t[p] - '0'); ++p; ++digits; }
  if (p < n && t[p] == '.') {
    ++p;
    while (p < n && t[p] >= '0' && t[p] <= '9') { m = m * 10.0 + (t[p] - '0'); ++p; ++digits; --scale; }
  }
  if (!digits) ok = false;
  if (p < n && (t[p] == 'e' || t[p] == 'E')) {
    ++p;
    bool eneg = false;
    if (p < n && (t[p] == '-' || t[p] == '+')) { eneg = t[p] == '-'; ++p; }
    int e = 0, ed = 0;
    while (p < n && t[p] >= '0' && t[p] <= '9') { e = min(e * 10 + (t[p] - '0'), 400); ++p; ++ed; }
    if (!ed) ok = false;
    scale += eneg ? -e : e;
  }
  double v = m;
  if (scale > 0) v = m * pow(10.0, double(scale));
  else if (scale < 0) v = m / pow(10.0, double(-scale));
  return neg ? -v : v;
}

// mode 0: rating line  "uid \t iid [\t rating [\t anything]]"  -> u, i, r (rating 1.0 when the column is missing)
// mode 1: negatives line "(u,i) \t n1 \t n2 ..."               -> tok_count[line] = number of ids after the first tab
// mode 2: negatives line, second pass                           -> ids written at neg_ptr[line] ...
__global__ void __launch_bounds__(256)
tsv_parse_kernel(const char* __restrict__ text, int64_t n, const int64_t* __restrict__ base, int64_t max_lines, int mode,
                 int32_t* __restrict__ out_u, int32_t* __restrict__ out_i, float* __restrict__ out_r,
                 int64_t* __restrict__ tok_count, const int64_t* __restrict__ neg_ptr, int32_t* __restrict__ neg_idx,
                 int32_t* __restrict__ err) {
  const int64_t b0 = int64_t(blockIdx.x) * kTsvBlock + threadIdx.x * 4;
  bool st[4];
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { st[k] = line_start_at(text, b0 + k, n); c += st[k] ? 1 : 0; }
  typedef cub::BlockScan<int, 256> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int prefix;
  Scan(tmp).ExclusiveSum(c, prefix);
  int64_t line = base[blockIdx.x] + prefix;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!st[k]) continue;
    const int64_t ln = line++;
    if (ln >= max_lines) { atomicOr(err, 2); continue; }
    int64_t p = b0 + k;
    bool ok = true;
    if (mode == 0) {
      const int64_t u = parse_uint(text, p, n, ok);
      if (!(p < n && text[p] == '\t')) ok = false;
      ++p;
      const int64_t i = parse_uint(text, p, n, ok);
      float r = 1.0f;
      if (p < n && text[p] == '\t') { ++p; r = float(parse_float(text, p, n, ok)); }
      out_u[ln] = int32_t(u); out_i[ln] = int32_t(i); out_r[ln] = r;
    } else {
      while (p < n && text[p] != '\t' && text[p] != '\n') ++p;      // the "(u,i)" field
      int64_t cnt = 0;
      const int64_t dst = mode == 2 ? neg_ptr[ln] : 0;
      while (p < n && text[p] == '\t') {
        ++p;
        if (p < n && (text[p] == '\n' || text[p] == '\r')) break;     // trailing tab
        const int64_t v = parse_uint(text, p, n, ok);
        if (mode == 2) neg_idx[dst + cnt] = int32_t(v);
        ++cnt;
      }
      if (mode == 1) tok_count[ln] = cnt;
    }
    if (!ok) atomicOr(err, 1);
  }
}

// trainList row of line k (Dataset.py:316-320): the cursor advances by at most one per line and never passes the line's
// uid.  For a uid-sorted file  row_k = min(k + 1, k + min_{j <= k}(uid_j - j)).
__global__ void __launch_bounds__(256) cursor_key_kernel(const int32_t* __restrict__ u, int64_t n, int64_t* __restrict__ key,
                                                         int32_t* __restrict__ unsorted) {
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x) {
    key[k] = int64_t(u[k]) - k;
    if (k > 0 && u[k] < u[k - 1]) atomicOr(unsorted, 1);
  }
}
__global__ void __launch_bounds__(256) cursor_row_kernel(const int64_t* __restrict__ pmin, const int32_t* __restrict__ u, int64_t n,
                                                         int quirk, int32_t* __restrict__ row) {
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x)
    row[k] = quirk ? int32_t(min(k + 1, k + pmin[k])) : u[k];
}

__global__ void __launch_bounds__(256) make_keys_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ item,
                                                        const float* __restrict__ rating, int64_t n, uint64_t* __restrict__ key,
                                                        int64_t* __restrict__ index) {
  // rating == nullptr: every line counts; else lines with rating <= 0 get the all-ones key (sorted last, dropped)
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x) {
    const bool keep = !rating || rating[k] > 0.f;
    key[k] = keep ? (uint64_t(uint32_t(row[k])) << 32) | uint32_t(item[k]) : ~uint64_t(0);
    if (index) index[k] = k;
  }
}
// flag[k] = 1 iff sorted key k is the first of its group (and not the dropped key)
__global__ void __launch_bounds__(256) first_flags_kernel(const uint64_t* __restrict__ key, int64_t n, unsigned char* __restrict__ flag) {
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x)
    flag[k] = (key[k] != ~uint64_t(0) && (k == 0 || key[k] != key[k - 1])) ? 1 : 0;
}
__global__ void __launch_bounds__(256) csr_fill_kernel(const uint64_t* __restrict__ ukey, const int64_t* __restrict__ n_unique, int64_t rows,
                                                       int64_t* __restrict__ ptr, int32_t* __restrict__ idx) {
  const int64_t m = *n_unique;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = int64_t(ukey[k] >> 32);
    idx[k] = int32_t(uint32_t(ukey[k]));
    // ptr[r'] = k for every row r' in (row of the previous key, r]
    const int64_t rp = k == 0 ? -1 : int64_t(ukey[k - 1] >> 32);
    for (int64_t q = rp + 1; q <= r && q <= rows; ++q) ptr[q] = k;
    if (k == m - 1) for (int64_t q = r + 1; q <= rows; ++q) ptr[q] = m;
  }
  if (m == 0) for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q <= rows; q += int64_t(gridDim.x) * blockDim.x) ptr[q] = 0;
}
__global__ void __launch_bounds__(256) gather_pairs_kernel(const int64_t* __restrict__ order, const int64_t* __restrict__ n_sel,
                                                           const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                                           int32_t* __restrict__ out_u, int32_t* __restrict__ out_i) {
  const int64_t m = *n_sel;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
    out_u[k] = u[order[k]];
    out_i[k] = i[order[k]];
  }
}

__global__ void __launch_bounds__(256) pad_tail_kernel(int64_t* __restrict__ sel, const int64_t* __restrict__ m, int64_t n) {
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x)
    if (k >= *m) sel[k] = int64_t(0x7fffffffffffffffLL);
}

struct MinOp { __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a < b ? a : b; } };

static inline int grid1d(int64_t n) {
  return int(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, int64_t(sm_count()) * 16)));
}
// bump allocator over the caller's workspace
struct Arena {
  char* base; int64_t size, off;
  template <typename T> T* take(int64_t count) {
    off = (off + 255) & ~int64_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * int64_t(sizeof(T));
    return off <= size ? p : nullptr;
  }
};

}  // namespace apr

using namespace apr;

extern "C" {

int64_t apr_loader_workspace_bytes(int64_t n_bytes, int64_t max_lines) {
  if (n_bytes < 0 || max_lines < 0) return -1;
  // block counts + bases, four 8-byte arrays and a flag array per line, CUB temporaries (radix sort of 16-byte pairs
  // needs about one extra copy of its input; scans a few KB)
  return 2 * 8 * (n_bytes / kTsvBlock + 2) + max_lines * (8 * 6 + 1) + (int64_t(32) << 20) + 4096;
}

/* Line count of a TSV text (device buffer): n_lines_host receives the number of non-empty lines.  Synchronises. */
int apr_tsv_count_lines(const char* text, int64_t n_bytes, void* ws, int64_t ws_bytes, int64_t* n_lines_host,
                        apr_stream_t stream) {
  if (!text || n_bytes < 1 || !ws || !n_lines_host) return APR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nblk = (n_bytes + kTsvBlock - 1) / kTsvBlock;
  Arena a{static_cast<char*>(ws), ws_bytes, 0};
  int64_t* counts = a.take<int64_t>(nblk + 1);
  int64_t* base = a.take<int64_t>(nblk + 1);
  if (!counts || !base) return APR_E_WORKSPACE;
  APR_CUDA_CHECK(cudaMemsetAsync(counts + nblk, 0, 8, st));
  tsv_count_kernel<<<unsigned(nblk), 256, 0, st>>>(text, n_bytes, counts);
  APR_LAUNCH_CHECK();
  size_t tb = 0;
  APR_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts, base, int(nblk + 1), st));
  void* tmp = a.take<char>(int64_t(tb));
  if (!tmp) return APR_E_WORKSPACE;
  APR_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, counts, base, int(nblk + 1), st));
  APR_CUDA_CHECK(cudaMemcpyAsync(n_lines_host, base + nblk, 8, cudaMemcpyDeviceToHost, st));
  APR_CUDA_CHECK(cudaStreamSynchronize(st));
  return APR_OK;
}

/* Parse (after apr_tsv_count_lines on the same workspace: the block bases are reused).
 * mode 0: rating file -> out_u / out_i / out_r [max_lines].
 * mode 1: negatives file, pass 1 -> tok_count [max_lines] (ids per line).
 * mode 2: negatives file, pass 2 -> neg_idx at neg_ptr[line].
 * *err_flag (device int32): bit 0 = malformed field, bit 1 = more lines than max_lines. */
int apr_tsv_parse(const char* text, int64_t n_bytes, const void* ws, int64_t max_lines, int32_t mode, int32_t* out_u,
                  int32_t* out_i, float* out_r, int64_t* tok_count, const int64_t* neg_ptr, int32_t* neg_idx,
                  int32_t* err_flag, apr_stream_t stream) {
  if (!text || n_bytes < 1 || !ws || !err_flag || mode < 0 || mode > 2) return APR_E_ARG;
  if (mode == 0 && (!out_u || !out_i || !out_r)) return APR_E_ARG;
  if (mode == 1 && !tok_count) return APR_E_ARG;
  if (mode == 2 && (!neg_ptr || !neg_idx)) return APR_E_ARG;
  const int64_t nblk = (n_bytes + kTsvBlock - 1) / kTsvBlock;
  Arena a{const_cast<char*>(static_cast<const char*>(ws)), int64_t(1) << 62, 0};
  a.take<int64_t>(nblk + 1);
  const int64_t* base = a.take<int64_t>(nblk + 1);
  tsv_parse_kernel<<<unsigned(nblk), 256, 0, static_cast<cudaStream_t>(stream)>>>(text, n_bytes, base, max_lines, mode, out_u, out_i,
                                                                                 out_r, tok_count, neg_ptr, neg_idx, err_flag);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

/* trainList row of every line (Dataset.py:306-325).  quirk != 0 reproduces the reference's cursor for a
 * uid-sorted file; *unsorted_flag (device int32) is set if the file is not sorted by uid -- the caller then falls back
 * to the sequential host cursor.  quirk == 0: row = uid. */
int apr_loader_train_rows(const int32_t* u, int64_t n, int32_t quirk, int32_t* row, int32_t* unsorted_flag, void* ws,
                          int64_t ws_bytes, apr_stream_t stream) {
  if (!u || n < 1 || !row || !unsorted_flag || !ws) return APR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a{static_cast<char*>(ws), ws_bytes, 0};
  int64_t* key = a.take<int64_t>(n);
  int64_t* pmin = a.take<int64_t>(n);
  if (!key || !pmin) return APR_E_WORKSPACE;
  cursor_key_kernel<<<grid1d(n), 256, 0, st>>>(u, n, key, unsorted_flag);
  size_t tb = 0;
  APR_CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, tb, key, pmin, MinOp(), int(n), st));
  void* tmp = a.take<char>(int64_t(tb));
  if (!tmp) return APR_E_WORKSPACE;
  APR_CUDA_CHECK(cub::DeviceScan::InclusiveScan(tmp, tb, key, pmin, MinOp(), int(n), st));
  cursor_row_kernel<<<grid1d(n), 256, 0, st>>>(pmin, u, n, quirk, row);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

/* Sorted, de-duplicated CSR of (row, item) (ptr int64 [rows + 1], idx int32 [<= n]); *n_unique_dev (device int64) = entries. */
int apr_loader_csr(const int32_t* row, const int32_t* item, int64_t n, int64_t rows, int64_t* ptr, int32_t* idx,
                   int64_t* n_unique_dev, void* ws, int64_t ws_bytes, apr_stream_t stream) {
  if (!row || !item || n < 1 || rows < 1 || !ptr || !idx || !n_unique_dev || !ws) return APR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a{static_cast<char*>(ws), ws_bytes, 0};
  uint64_t* k0 = a.take<uint64_t>(n);
  uint64_t* k1 = a.take<uint64_t>(n);
  uint64_t* uk = a.take<uint64_t>(n);
  if (!k0 || !k1 || !uk) return APR_E_WORKSPACE;
  make_keys_kernel<<<grid1d(n), 256, 0, st>>>(row, item, nullptr, n, k0, nullptr);
  size_t tb = 0, tb2 = 0;
  APR_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tb, k0, k1, int(n), 0, 64, st));
  APR_CUDA_CHECK(cub::DeviceSelect::Unique(nullptr, tb2, k1, uk, n_unique_dev, int(n), st));
  void* tmp = a.take<char>(int64_t(std::max(tb, tb2)));
  if (!tmp) return APR_E_WORKSPACE;
  APR_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp, tb, k0, k1, int(n), 0, 64, st));
  APR_CUDA_CHECK(cub::DeviceSelect::Unique(tmp, tb2, k1, uk, n_unique_dev, int(n), st));
  csr_fill_kernel<<<grid1d(n), 256, 0, st>>>(uk, n_unique_dev, rows, ptr, idx);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

/* The keys of the reference's dok trainMatrix in insertion order (Dataset.py:278-304 -> APR.py:30-36): lines with
 * rating > 0, duplicates of a (u, i) pair collapsed onto its first occurrence.  out_u / out_i [<= n]; *n_pairs_dev. */
int apr_loader_unique_pairs(const int32_t* u, const int32_t* i, const float* rating, int64_t n, int32_t* out_u, int32_t* out_i,
                            int64_t* n_pairs_dev, void* ws, int64_t ws_bytes, apr_stream_t stream) {
  if (!u || !i || !rating || n < 1 || !out_u || !out_i || !n_pairs_dev || !ws) return APR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a{static_cast<char*>(ws), ws_bytes, 0};
  uint64_t* k0 = a.take<uint64_t>(n);
  uint64_t* k1 = a.take<uint64_t>(n);
  int64_t* v0 = a.take<int64_t>(n);
  int64_t* v1 = a.take<int64_t>(n);
  int64_t* sel = a.take<int64_t>(n);
  unsigned char* flag = a.take<unsigned char>(n);
  if (!k0 || !k1 || !v0 || !v1 || !sel || !flag) return APR_E_WORKSPACE;
  make_keys_kernel<<<grid1d(n), 256, 0, st>>>(u, i, rating, n, k0, v0);
  size_t tb = 0, tb2 = 0, tb3 = 0;
  APR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, int(n), 0, 64, st));   // stable: first occurrence first
  APR_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, tb2, v1, flag, sel, n_pairs_dev, int(n), st));
  APR_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tb3, sel, v0, int(n), 0, 64, st));
  void* tmp = a.take<char>(int64_t(std::max(tb, std::max(tb2, tb3))));
  if (!tmp) return APR_E_WORKSPACE;
  APR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, int(n), 0, 64, st));
  first_flags_kernel<<<grid1d(n), 256, 0, st>>>(k1, n, flag);
  APR_CUDA_CHECK(cub::DeviceSelect::Flagged(tmp, tb2, v1, flag, sel, n_pairs_dev, int(n), st));
  // back to file order: the selected line indices ascending.  The tail of `sel` beyond *n_pairs_dev is padded with the
  // largest index so that a fixed-size sort leaves the selected ones in front.
  pad_tail_kernel<<<grid1d(n), 256, 0, st>>>(sel, n_pairs_dev, n);
  APR_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp, tb3, sel, v0, int(n), 0, 64, st));
  gather_pairs_kernel<<<grid1d(n), 256, 0, st>>>(v0, n_pairs_dev, u, i, out_u, out_i);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

}  // extern "C"
