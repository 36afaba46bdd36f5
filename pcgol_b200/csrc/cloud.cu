// cloud.cu — device-resident pc.PointCloud and its PCD encodings.
//
// pc.Unmarshal / pc.Marshal (pc/io.go:32-45,137-285): the header is text and is parsed on the host
// with the reference's rules; the payload is the record buffer of pc.PointCloud.  `DATA binary`
// goes to HBM verbatim, `DATA binary_compressed` is LZF-decoded on the host (a sequential format)
// and its field-major (SoA) layout is transposed into records on the device.  The resulting
// pcg_cloud feeds VoxelGrid, the index build and ICP without ever returning to the host, and
// Marshal reads it back once.
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "bvh.cuh"

namespace pcg {

struct CloudHeader {  // pc.PointCloudHeader (pc/pointcloud.go:9-18)
  float version = 0.f;
  std::vector<std::string> fields, type;
  std::vector<int64_t> size, count;
  int64_t width = 0, height = 0;
  std::vector<float> viewpoint;
  int64_t stride() const {  // pc/pointcloud.go:64-70
    int64_t s = 0;
    for (size_t i = 0; i < size.size() && i < count.size(); i++) s += size[i] * count[i];
    return s;
  }
};

struct Cloud {
  int device = 0;
  CloudHeader h;
  int64_t points = 0;
  int64_t bytes = 0;  // len(Data): points*stride, except after a binary_compressed load with slack (io.go:217)
  uint8_t* d_data = nullptr;
};

void cloud_free(Cloud* c) {
  if (!c) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  if (c->d_data) cudaFree(c->d_data);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
}

// ---- host-side text rules of the reference ------------------------------------------------
static std::vector<std::string> go_fields(const std::string& line) {  // strings.Fields
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && isspace((unsigned char)line[i])) i++;
    size_t j = i;
    while (j < line.size() && !isspace((unsigned char)line[j])) j++;
    if (j > i) out.push_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}
static int64_t go_atoi(const std::string& s) {  // strconv.Atoi
  errno = 0;
  char* end = nullptr;
  bool ok = !s.empty() && (isdigit((unsigned char)s[0]) || ((s[0] == '+' || s[0] == '-') && s.size() > 1));
  long long v = ok ? strtoll(s.c_str(), &end, 10) : 0;
  if (!ok || *end != 0 || errno == ERANGE)
    throw StatusError{PCG_E_PCD_SYNTAX, "strconv.Atoi: parsing \"" + s + "\": invalid syntax"};
  return v;
}
static float go_parse_f32(const std::string& s) {  // strconv.ParseFloat(s, 32)
  errno = 0;
  char* end = nullptr;
  bool ok = !s.empty() && !isspace((unsigned char)s[0]) && s.find('_') == std::string::npos;
  float v = ok ? strtof(s.c_str(), &end) : 0.f;
  if (!ok || end == s.c_str() || *end != 0 || (errno == ERANGE && std::isinf(v)))
    throw StatusError{PCG_E_PCD_SYNTAX, "strconv.ParseFloat: parsing \"" + s + "\": invalid syntax"};
  return v;
}
static uint32_t go_parse_u32(const std::string& s) {  // strconv.ParseUint(s, 10, 32)
  bool ok = !s.empty() && s.size() <= 10;
  unsigned long long v = 0;
  for (char ch : s) {
    if (!isdigit((unsigned char)ch)) ok = false;
    v = v * 10 + (unsigned)(ch - '0');
  }
  if (!ok || v > 0xffffffffull)
    throw StatusError{PCG_E_PCD_SYNTAX, "strconv.ParseUint: parsing \"" + s + "\": invalid syntax"};
  return (uint32_t)v;
}

struct ByteReader {  // bufio.Reader over the caller's bytes
  const uint8_t* b;
  int64_t n, pos = 0;
  std::string read_line() {  // ReadLine: io.EOF when nothing is left; strips \n / \r\n
    if (pos >= n) throw StatusError{PCG_E_PCD_EOF, "EOF"};
    int64_t end = pos;
    while (end < n && b[end] != '\n') end++;
    std::string line((const char*)b + pos, (size_t)(end - pos));
    pos = end < n ? end + 1 : n;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return line;
  }
  const uint8_t* read_full(int64_t k) {  // io.ReadFull
    if (pos + k > n) throw StatusError{PCG_E_PCD_EOF, pos < n ? "unexpected EOF" : "EOF"};
    const uint8_t* p = b + pos;
    pos += k;
    return p;
  }
  int32_t read_i32() {  // binary.Read(rb, LittleEndian, &int32)
    const uint8_t* p = read_full(4);
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
  }
};

enum { kFmtAscii = 0, kFmtBinary = 1, kFmtCompressed = 2 };

// unmarshalPCDHeaderTo (io.go:47-135)
static void parse_header(ByteReader& r, CloudHeader& h, int64_t* n_points, int* fmt) {
  *n_points = 0;
  for (;;) {
    const std::vector<std::string> a = go_fields(r.read_line());
    if (a.size() < 2) throw StatusError{PCG_E_PCD_SYNTAX, "header field must have value"};
    const std::string& k = a[0];
    if (k == "VERSION") {
      h.version = go_parse_f32(a[1]);
    } else if (k == "FIELDS") {
      h.fields.assign(a.begin() + 1, a.end());
    } else if (k == "SIZE") {
      h.size.clear();
      for (size_t i = 1; i < a.size(); i++) h.size.push_back(go_atoi(a[i]));
    } else if (k == "TYPE") {
      h.type.assign(a.begin() + 1, a.end());
    } else if (k == "COUNT") {
      h.count.clear();
      for (size_t i = 1; i < a.size(); i++) h.count.push_back(go_atoi(a[i]));
    } else if (k == "WIDTH") {
      h.width = go_atoi(a[1]);
    } else if (k == "HEIGHT") {
      h.height = go_atoi(a[1]);
    } else if (k == "VIEWPOINT") {
      h.viewpoint.clear();
      for (size_t i = 1; i < a.size(); i++) h.viewpoint.push_back(go_parse_f32(a[i]));
    } else if (k == "POINTS") {
      *n_points = go_atoi(a[1]);
    } else if (k == "DATA") {
      if (a[1] == "ascii")
        *fmt = kFmtAscii;
      else if (a[1] == "binary")
        *fmt = kFmtBinary;
      else if (a[1] == "binary_compressed")
        *fmt = kFmtCompressed;
      else
        throw StatusError{PCG_E_PCD_SYNTAX, "unknown data format"};
      break;
    }
  }
  if (h.fields.size() != h.size.size()) throw StatusError{PCG_E_PCD_SYNTAX, "size field size is wrong"};
  if (h.fields.size() != h.type.size()) throw StatusError{PCG_E_PCD_SYNTAX, "type field size is wrong"};
  if (h.fields.size() != h.count.size()) throw StatusError{PCG_E_PCD_SYNTAX, "count field size is wrong"};
}

// LZF (liblzf format; the reference calls github.com/zhuyie/golzf, go.mod:5, a port of liblzf).
static int64_t lzf_decompress(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_len) {
  int64_t ip = 0, op = 0;
  while (ip < in_len) {
    uint32_t ctrl = in[ip++];
    if (ctrl < 32) {
      ctrl++;
      if (op + ctrl > out_len) throw StatusError{PCG_E_PCD_CORRUPT, "lzf: insufficient buffer"};
      if (ip + ctrl > in_len) throw StatusError{PCG_E_PCD_CORRUPT, "lzf: data corruption"};
      memcpy(out + op, in + ip, ctrl);
      ip += ctrl;
      op += ctrl;
    } else {
      int64_t len = ctrl >> 5;
      int64_t ref = op - (int64_t)((ctrl & 0x1f) << 8) - 1;
      if (ip >= in_len) throw StatusError{PCG_E_PCD_CORRUPT, "lzf: data corruption"};
      if (len == 7) {
        len += in[ip++];
        if (ip >= in_len) throw StatusError{PCG_E_PCD_CORRUPT, "lzf: data corruption"};
      }
      ref -= in[ip++];
      if (op + len + 2 > out_len) throw StatusError{PCG_E_PCD_CORRUPT, "lzf: insufficient buffer"};
      if (ref < 0) throw StatusError{PCG_E_PCD_CORRUPT, "lzf: data corruption"};
      for (int64_t k = 0; k < len + 2; k++) out[op++] = out[ref++];
    }
  }
  return op;
}

// io.go:205-228: dec holds field after field (SoA); record p, field i <- dec[head[i] + p*size[i] ..][size[i]]
// (sic: the source advances by size, not size*count, and copies `size` bytes - fields with COUNT > 1 keep only
// their first element; restated as is).  One thread per (point, field).
constexpr int kMaxFields = 32;
struct SoaLayout {
  int32_t n_fields;
  int64_t head[kMaxFields];
  int32_t offset[kMaxFields];
  int32_t size[kMaxFields];
};
__global__ void __launch_bounds__(256)
    pcd_soa_to_aos_kernel(const uint8_t* __restrict__ dec, uint8_t* __restrict__ data, int64_t points, int64_t stride,
                          SoaLayout L) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = t / L.n_fields;
  const int i = (int)(t - p * L.n_fields);
  if (p >= points) return;
  const int size = L.size[i];
  const uint8_t* src = dec + L.head[i] + p * size;
  uint8_t* dst = data + p * stride + L.offset[i];
  if (size == 4 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 3) == 0) {
    *(uint32_t*)dst = *(const uint32_t*)src;
  } else {
    for (int b = 0; b < size; b++) dst[b] = src[b];
  }
}

Cloud* cloud_new_device(const CloudHeader& h, int64_t points, int64_t bytes, int device, cudaStream_t stream) {
  if (h.fields.size() > (size_t)kMaxFields) throw StatusError{PCG_E_TOO_LARGE, "more than 32 fields"};
  Cloud* c = new Cloud();
  c->device = device;
  c->h = h;
  c->points = points;
  c->bytes = bytes;
  try {
    PCG_CUDA(cudaMallocAsync((void**)&c->d_data, (size_t)std::max<int64_t>(1, bytes), stream));
  } catch (...) {
    delete c;
    throw;
  }
  return c;
}

// pc.Unmarshal (io.go:32-45)
Cloud* cloud_unmarshal(const uint8_t* pcd, int64_t len, int device, cudaStream_t stream) {
  ByteReader r{pcd, len};
  CloudHeader h;
  int64_t points = 0;
  int fmt = kFmtBinary;
  parse_header(r, h, &points, &fmt);
  for (size_t i = 0; i < h.size.size(); i++)
    if (h.size[i] < 0 || h.count[i] < 0 || h.size[i] > (1 << 20) || h.count[i] > (1 << 20))
      throw StatusError{PCG_E_REF_WOULD_PANIC, "negative or absurd SIZE / COUNT"};
  const int64_t stride = h.stride();
  if (points < 0 || (stride > 0 && points > ((int64_t)1 << 40) / stride))
    throw StatusError{PCG_E_REF_WOULD_PANIC, "makeslice: len out of range"};
  if (fmt == kFmtBinary) {  // io.go:181-186
    const uint8_t* payload = r.read_full(points * stride);
    Cloud* c = cloud_new_device(h, points, points * stride, device, stream);
    if (c->bytes) PCG_CUDA(cudaMemcpyAsync(c->d_data, payload, (size_t)c->bytes, cudaMemcpyHostToDevice, stream));
    PCG_CUDA(cudaStreamSynchronize(stream));  // the caller's bytes are not referenced after return
    return c;
  }
  if (fmt == kFmtAscii) {  // io.go:140-180
    // every record takes at least one byte of text per F/U value (plus a separator): a POINTS count that the input
    // cannot possibly hold would only make the reference allocate and zero a huge slice - refuse it instead of
    // following it (a tiny file must not be able to demand a terabyte)
    if (points * stride > (int64_t)64 * (len + 1024))
      throw StatusError{PCG_E_TOO_LARGE, "ascii PCD: POINTS * record size is out of proportion to the input length"};
    std::vector<uint8_t> data((size_t)(points * stride), 0);
    int64_t off = 0;
    while (r.pos < r.n) {
      const std::vector<std::string> tok = go_fields(r.read_line());
      size_t lo = 0;
      for (size_t i = 0; i < h.type.size(); i++) {
        for (int64_t j = 0; j < h.count[i]; j++) {
          const bool isf = h.type[i] == "F", isu = h.type[i] == "U";
          if (isf || isu) {
            if (lo + (size_t)j >= tok.size()) throw StatusError{PCG_E_REF_WOULD_PANIC, "index out of range (too few values on a line)"};
            uint32_t bits;
            if (isf) {
              const float v = go_parse_f32(tok[lo + (size_t)j]);
              memcpy(&bits, &v, 4);
            } else {
              bits = go_parse_u32(tok[lo + (size_t)j]);
            }
            if (off + 4 > (int64_t)data.size()) throw StatusError{PCG_E_REF_WOULD_PANIC, "slice bounds out of range (more values than POINTS)"};
            data[(size_t)off] = (uint8_t)bits;
            data[(size_t)off + 1] = (uint8_t)(bits >> 8);
            data[(size_t)off + 2] = (uint8_t)(bits >> 16);
            data[(size_t)off + 3] = (uint8_t)(bits >> 24);
          }
          off += h.size[i];
        }
        lo += (size_t)h.count[i];
      }
    }
    Cloud* c = cloud_new_device(h, points, (int64_t)data.size(), device, stream);
    if (c->bytes) PCG_CUDA(cudaMemcpyAsync(c->d_data, data.data(), data.size(), cudaMemcpyHostToDevice, stream));
    PCG_CUDA(cudaStreamSynchronize(stream));
    return c;
  }
  // binary_compressed (io.go:187-228)
  const int32_t n_comp = r.read_i32();
  const int32_t n_unc = r.read_i32();
  if (n_comp < 0 || n_unc < 0) throw StatusError{PCG_E_REF_WOULD_PANIC, "makeslice: len out of range"};
  const uint8_t* comp = r.read_full(n_comp);
  std::vector<uint8_t> dec((size_t)n_unc);
  const int64_t got = lzf_decompress(comp, n_comp, dec.data(), n_unc);
  if (got != n_unc) throw StatusError{PCG_E_PCD_CORRUPT, "wrong uncompressed size"};
  if (h.fields.size() > (size_t)kMaxFields) throw StatusError{PCG_E_TOO_LARGE, "more than 32 fields"};
  SoaLayout L;
  L.n_fields = (int32_t)h.fields.size();
  int64_t pos = 0, off = 0;
  for (int i = 0; i < L.n_fields; i++) {
    L.head[i] = pos;
    L.offset[i] = (int32_t)off;
    L.size[i] = (int32_t)h.size[(size_t)i];
    pos += h.size[(size_t)i] * h.count[(size_t)i] * points;
    off += h.size[(size_t)i] * h.count[(size_t)i];
    if (points > 0 && (L.head[i] + points * L.size[i] > n_unc || (points - 1) * stride + L.offset[i] + L.size[i] > n_unc))
      throw StatusError{PCG_E_REF_WOULD_PANIC, "slice bounds out of range (uncompressed size smaller than the records)"};
  }
  Cloud* c = cloud_new_device(h, points, n_unc, device, stream);  // len(Data) == nUncompressed (io.go:217)
  try {
    if (n_unc) {
      DevBuf<uint8_t> d_dec((size_t)n_unc, stream);
      PCG_CUDA(cudaMemcpyAsync(d_dec.p, dec.data(), (size_t)n_unc, cudaMemcpyHostToDevice, stream));
      PCG_CUDA(cudaMemsetAsync(c->d_data, 0, (size_t)n_unc, stream));
      if (points > 0 && L.n_fields > 0)
        PCG_LAUNCH(pcd_soa_to_aos_kernel, div_up(points * L.n_fields, 256), 256, 0, stream, d_dec.p, c->d_data, points,
                   stride, L);
      PCG_CUDA(cudaStreamSynchronize(stream));
    }
  } catch (...) {
    cloud_free(c);
    throw;
  }
  return c;
}

// pc.Marshal header (io.go:232-277)
std::string cloud_marshal_header(const Cloud& c) {
  const CloudHeader& h = c.h;
  std::vector<float> vp = h.viewpoint;
  if (vp.empty()) vp = {0, 0, 0, 1, 0, 0, 0};  // io.go:248-250
  char buf[64];
  std::string s;
  snprintf(buf, sizeof(buf), "VERSION %0.1f\n", (double)h.version);
  s += buf;
  auto join = [](const std::vector<std::string>& v) {
    std::string o;
    for (size_t i = 0; i < v.size(); i++) o += (i ? " " : "") + v[i];
    return o;
  };
  auto join_i = [](const std::vector<int64_t>& v) {
    std::string o;
    for (size_t i = 0; i < v.size(); i++) o += (i ? " " : "") + std::to_string(v[i]);
    return o;
  };
  s += "FIELDS " + join(h.fields) + "\nSIZE " + join_i(h.size) + "\nTYPE " + join(h.type) + "\nCOUNT " + join_i(h.count);
  s += "\nWIDTH " + std::to_string(h.width) + "\nHEIGHT " + std::to_string(h.height) + "\nVIEWPOINT ";
  for (size_t i = 0; i < vp.size(); i++) {
    snprintf(buf, sizeof(buf), "%s%.4f", i ? " " : "", (double)vp[i]);  // FormatFloat(v, 'f', 4, 32)
    s += buf;
  }
  s += "\nPOINTS " + std::to_string(c.points) + "\nDATA binary\n";
  return s;
}

// PointCloud.Vec3Iterator field resolution (pointcloud.go:130-171): "xyz" if it comes first, else consecutive
// x,y,z, else the three fields wherever they are.  Returns false when a coordinate field is missing.
bool cloud_xyz_offsets(const CloudHeader& h, int64_t off_out[3]) {
  auto field_off = [&](const std::string& name, int64_t* out) {
    int64_t off = 0;
    for (size_t i = 0; i < h.fields.size(); i++) {
      if (h.fields[i] == name) {
        *out = off;
        return true;
      }
      off += h.size[i] * h.count[i];
    }
    return false;
  };
  int xyz = 0;
  std::string first;
  for (const std::string& name : h.fields) {
    if (name == "xyz") {
      xyz = 3;
      first = name;
      break;
    }
    if (name == "x" && xyz == 0) {
      xyz = 1;
      first = name;
    } else if (name == "y" && xyz == 1) {
      xyz = 2;
    } else if (name == "z" && xyz == 2) {
      xyz = 3;
      break;
    } else {
      xyz = 0;
    }
  }
  if (xyz == 3) {
    int64_t o;
    if (!field_off(first, &o)) return false;
    if ((h.stride() & 3) == 0 && (o & 3) == 0) {  // the aligned float32Iterator doubles as the Vec3Iterator
      off_out[0] = o;
      off_out[1] = o + 4;
      off_out[2] = o + 8;
      return true;
    }
  }
  return field_off("x", &off_out[0]) && field_off("y", &off_out[1]) && field_off("z", &off_out[2]);
}

}  // namespace pcg

// ---- C ABI ---------------------------------------------------------------------------------
// (kept in this translation unit: it needs the full Cloud type)
namespace pcg {
pcg_status voxelgrid_filter_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], uint8_t* d_out,
                                   int64_t* n_out, cudaStream_t stream);
pcg_status icp_fit_device(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, bool evaluate_only,
                          float trans[16], pcg_icp_stat* stat, cudaStream_t stream);
pcg_status api_guard(const std::function<pcg_status()>& f);
pcg_index* api_wrap_index(Index* ix);
Index* api_index_of(pcg_index* idx);
void api_check_device(int device);
}  // namespace pcg
using namespace pcg;

struct pcg_cloud {
  Cloud* c;
};

static CloudView cloud_view(const Cloud& c) {
  int64_t off[3];
  if (!cloud_xyz_offsets(c.h, off)) throw StatusError{PCG_E_INVALID_FIELD, "invalid field name"};
  const int64_t stride = c.h.stride();
  check_view_args(c.d_data, c.points, stride, off);
  // binary_compressed keeps len(Data) == nUncompressed (io.go:217), which a crafted header can make smaller than
  // POINTS records: the reference would panic on the slice, here every consumer would read past the allocation
  if (c.points * stride > c.bytes)
    throw StatusError{PCG_E_REF_WOULD_PANIC, "slice bounds out of range (Data is shorter than POINTS records)"};
  return make_view(c.d_data, c.points, stride, off);
}

extern "C" {

pcg_status pcg_pcd_unmarshal(const void* pcd, int64_t len, int32_t device, pcg_cloud** out) {
  return api_guard([&]() -> pcg_status {
    if (!out || len < 0 || (len && !pcd)) throw StatusError{PCG_E_INVALID_ARG, "bad arguments"};
    *out = nullptr;
    api_check_device(device);
    DeviceGuard g(device);
    Cloud* c = cloud_unmarshal((const uint8_t*)pcd, len, device, cudaStreamPerThread);
    *out = new pcg_cloud{c};
    return PCG_OK;
  });
}

pcg_status pcg_pcd_marshal(const pcg_cloud* c, void* buf, int64_t cap, int64_t* len) {
  return api_guard([&]() -> pcg_status {
    if (!c || !len) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    const std::string head = cloud_marshal_header(*c->c);
    *len = (int64_t)head.size() + c->c->bytes;
    if (!buf || cap < *len) throw StatusError{PCG_E_INVALID_ARG, "buffer too small (len holds the size needed)"};
    memcpy(buf, head.data(), head.size());
    if (c->c->bytes) {
      DeviceGuard g(c->c->device);
      PCG_CUDA(cudaMemcpy((uint8_t*)buf + head.size(), c->c->d_data, (size_t)c->c->bytes, cudaMemcpyDeviceToHost));
    }
    return PCG_OK;
  });
}

pcg_status pcg_cloud_upload(const pcg_cloud_header* hd, const void* data, int32_t device, pcg_cloud** out) {
  return api_guard([&]() -> pcg_status {
    if (!hd || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    if (hd->n_fields < 0 || hd->n_fields > PCG_MAX_FIELDS || hd->n_viewpoint < 0 || hd->n_viewpoint > 16 ||
        hd->points < 0)
      throw StatusError{PCG_E_INVALID_ARG, "bad header"};
    CloudHeader h;
    h.version = hd->version;
    for (int i = 0; i < hd->n_fields; i++) {
      h.fields.push_back(std::string(hd->fields[i], strnlen(hd->fields[i], sizeof(hd->fields[i]))));
      h.type.push_back(std::string(hd->type[i], strnlen(hd->type[i], sizeof(hd->type[i]))));
      if (hd->size[i] < 0 || hd->count[i] < 0) throw StatusError{PCG_E_INVALID_ARG, "negative SIZE / COUNT"};
      h.size.push_back(hd->size[i]);
      h.count.push_back(hd->count[i]);
    }
    h.width = hd->width;
    h.height = hd->height;
    h.viewpoint.assign(hd->viewpoint, hd->viewpoint + hd->n_viewpoint);
    const int64_t bytes = hd->points * h.stride();
    if (bytes && !data) throw StatusError{PCG_E_INVALID_ARG, "null data"};
    api_check_device(device);
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    Cloud* c = cloud_new_device(h, hd->points, bytes, device, s);
    try {
      if (bytes) PCG_CUDA(cudaMemcpyAsync(c->d_data, data, (size_t)bytes, cudaMemcpyHostToDevice, s));
      PCG_CUDA(cudaStreamSynchronize(s));
    } catch (...) {
      cloud_free(c);
      throw;
    }
    *out = new pcg_cloud{c};
    return PCG_OK;
  });
}

pcg_status pcg_cloud_get_header(const pcg_cloud* c, pcg_cloud_header* out) {
  return api_guard([&]() -> pcg_status {
    if (!c || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    const CloudHeader& h = c->c->h;
    memset(out, 0, sizeof(*out));
    if (h.fields.size() > PCG_MAX_FIELDS || h.viewpoint.size() > 16)
      throw StatusError{PCG_E_TOO_LARGE, "header does not fit pcg_cloud_header"};
    out->version = h.version;
    out->n_fields = (int32_t)h.fields.size();
    for (size_t i = 0; i < h.fields.size(); i++) {
      snprintf(out->fields[i], sizeof(out->fields[i]), "%s", h.fields[i].c_str());
      snprintf(out->type[i], sizeof(out->type[i]), "%s", h.type[i].c_str());
      out->size[i] = h.size[i];
      out->count[i] = h.count[i];
    }
    out->width = h.width;
    out->height = h.height;
    out->n_viewpoint = (int32_t)h.viewpoint.size();
    for (size_t i = 0; i < h.viewpoint.size(); i++) out->viewpoint[i] = h.viewpoint[i];
    out->points = c->c->points;
    out->data_bytes = c->c->bytes;
    return PCG_OK;
  });
}

pcg_status pcg_cloud_download(const pcg_cloud* c, void* data, int64_t cap) {
  return api_guard([&]() -> pcg_status {
    if (!c || cap < c->c->bytes || (c->c->bytes && !data)) throw StatusError{PCG_E_INVALID_ARG, "buffer too small"};
    if (c->c->bytes == 0) return PCG_OK;
    DeviceGuard g(c->c->device);
    PCG_CUDA(cudaMemcpy(data, c->c->d_data, (size_t)c->c->bytes, cudaMemcpyDeviceToHost));
    return PCG_OK;
  });
}

const void* pcg_cloud_device_ptr(const pcg_cloud* c) { return c ? c->c->d_data : nullptr; }

void pcg_cloud_free(pcg_cloud* c) {
  if (!c) return;
  cloud_free(c->c);
  delete c;
}

pcg_status pcg_cloud_voxelgrid_filter(const pcg_cloud* in, const float leaf[3], const int64_t chunk[3],
                                      pcg_cloud** out) {
  return api_guard([&]() -> pcg_status {
    if (!in || !out || !leaf || !chunk) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    for (int k = 0; k < 3; k++)
      if (chunk[k] < 0) throw StatusError{PCG_E_INVALID_ARG, "negative chunk size"};
    const Cloud& ci = *in->c;
    DeviceGuard g(ci.device);
    cudaStream_t s = cudaStreamPerThread;
    const CloudView v = cloud_view(ci);
    Cloud* co = cloud_new_device(ci.h, 0, ci.points * v.stride, ci.device, s);
    try {
      int64_t n_out = 0;
      const pcg_status rc = voxelgrid_filter_device(v, leaf, chunk, co->d_data, &n_out, s);
      if (rc != PCG_OK) {
        cloud_free(co);
        return rc;
      }
      co->points = n_out;  // voxelgrid.go:119-128
      co->bytes = n_out * v.stride;
      co->h.width = n_out;
      co->h.height = 1;
    } catch (...) {
      cloud_free(co);
      throw;
    }
    *out = new pcg_cloud{co};
    return PCG_OK;
  });
}

pcg_status pcg_cloud_index_build(const pcg_cloud* c, pcg_index** out) {
  return api_guard([&]() -> pcg_status {
    if (!c || !out) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    *out = nullptr;
    DeviceGuard g(c->c->device);
    cudaStream_t s = cudaStreamPerThread;
    Index* ix = index_build_device(cloud_view(*c->c), c->c->device, s);
    PCG_CUDA(cudaStreamSynchronize(s));
    *out = api_wrap_index(ix);
    return PCG_OK;
  });
}

pcg_status pcg_cloud_icp_fit(pcg_index* base, const pcg_cloud* target, const pcg_icp_params* params, float trans[16],
                             pcg_icp_stat* stat) {
  return api_guard([&]() -> pcg_status {
    if (!base || !target || !params || !trans) throw StatusError{PCG_E_INVALID_ARG, "null argument"};
    Index* ix = api_index_of(base);
    if (ix->device != target->c->device) throw StatusError{PCG_E_INVALID_ARG, "index and cloud live on different devices"};
    DeviceGuard g(ix->device);
    const pcg_status rc = icp_fit_device(*ix, cloud_view(*target->c), *params, false, trans, stat, cudaStreamPerThread);
    if (rc == PCG_E_NOT_ENOUGH_PAIRS) set_error("not enough correspondence pairs");
    return rc;
  });
}

}  // extern "C"
