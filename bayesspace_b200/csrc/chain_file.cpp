// chain_file.cpp -- chain persistence: snapshots stream from the host layer to disk at every thin boundary instead
// of being materialised as R lists.
//
// Replaces the way /root/reference/R/mcmcChain.R:48-73 (.write_chain) and :181-223 (.clean_chain,
// .flatten_matrix_list) persist a chain: one HDF5 dataset per parameter (z, mu, lambda, weights, Y, Ychange, plogLik),
// one ROW per kept iteration, each row the parameter of that iteration flattened ROW-major (as.vector(t(x))), with a
// "dims" attribute carrying the per-iteration shape.  libhdf5 is not part of this image, so the container is a
// directory with one raw little-endian file per parameter ("<param>.bin", rows back to back) plus "meta.json"
// (dtype, rows, cols, dims, file per parameter): the same matrices, readable with readBin() in ten lines of R
//  or numpy.fromfile.
//
// The samplers hand rows to a registered sink (bsb200_set_snapshot_sink) as they are produced; with a sink the
// caller may pass NULL for z / weights / Y in the *_out structs, and spatialEnhance's 7.2 GB list of Y snapshots
// (BASELINE config C3) never exists in host memory.
#include <sys/stat.h>

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.hpp"

namespace bsb {

const bsb200_sink *g_sink = nullptr;
static bsb200_sink g_sink_copy;

int sink_write(const char *param, int64_t row, const void *data, int64_t count, int is_f64, const int64_t *dims, int ndims) {
  if (!g_sink || !g_sink->write) return BSB200_OK;
  const int rc = g_sink->write(g_sink->user, param, row, data, count, is_f64, dims, ndims);
  return rc ? fail(BSB200_ERR_BUFFER, "the snapshot sink failed on %s row %lld (code %d)", param, (long long) row, rc) : BSB200_OK;
}
bool sink_active() { return g_sink && g_sink->write; }

namespace {

struct ParamFile {
  FILE *f = nullptr;
  int is_f64 = 1;
  int64_t rows = 0, cols = 0;
  std::vector<int64_t> dims;
};

struct ChainFile {
  std::string dir;
  std::map<std::string, ParamFile> params;
  std::string error;
};
ChainFile *g_file = nullptr;

int file_write(void *user, const char *param, int64_t row, const void *data, int64_t count, int32_t is_f64, const int64_t *dims,
               int32_t ndims) {
  ChainFile *cf = (ChainFile *) user;
  ParamFile &p = cf->params[param];
  if (!p.f) {
    const std::string path = cf->dir + "/" + param + ".bin";
    p.f = std::fopen(path.c_str(), "wb");
    if (!p.f) { cf->error = "cannot create " + path + ": " + std::strerror(errno); return 1; }
    p.is_f64 = is_f64;
    p.cols = count;
    p.dims.assign(dims, dims + ndims);
  }
  if (count != p.cols || is_f64 != p.is_f64) { cf->error = std::string("row shape of ") + param + " changed"; return 2; }
  const size_t esz = is_f64 ? 8 : 4;
  if (row != p.rows) {   // rows normally arrive in order; a gap (rows the chain never produces) is zero-filled
    if (row < p.rows) { cf->error = std::string("row of ") + param + " written twice"; return 3; }
    std::vector<char> zeros((size_t) count * esz, 0);
    for (; p.rows < row; p.rows++)
      if (std::fwrite(zeros.data(), esz, (size_t) count, p.f) != (size_t) count) { cf->error = "short write"; return 4; }
  }
  if (std::fwrite(data, esz, (size_t) count, p.f) != (size_t) count) { cf->error = std::string("short write on ") + param; return 4; }
  p.rows++;
  return 0;
}

}   // namespace
}   // namespace bsb

using namespace bsb;

extern "C" {

void bsb200_set_snapshot_sink(const bsb200_sink *sink) {
  if (sink) { g_sink_copy = *sink; g_sink = &g_sink_copy; }
  else g_sink = nullptr;
}

int bsb200_chain_file_open(const char *dir) {
  if (!dir || !*dir) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (g_file) return fail(BSB200_ERR_INVALID, "a chain file is already open (%s)", g_file->dir.c_str());
  if (mkdir(dir, 0777) != 0 && errno != EEXIST) return fail(BSB200_ERR_BUFFER, "cannot create directory %s: %s", dir, std::strerror(errno));
  g_file = new ChainFile();
  g_file->dir = dir;
  bsb200_sink s;
  s.write = file_write;
  s.user = g_file;
  bsb200_set_snapshot_sink(&s);
  return BSB200_OK;
}

int bsb200_chain_file_close(void) {
  if (!g_file) return fail(BSB200_ERR_INVALID, "no chain file is open");
  bsb200_set_snapshot_sink(nullptr);
  ChainFile *cf = g_file;
  g_file = nullptr;
  int rc = BSB200_OK;
  std::string meta = "{\n \"format\": \"bsb200-chain-1\",\n \"layout\": \"one row per kept iteration, each row flattened row-major; little-endian\",\n \"params\": {";
  bool first = true;
  for (auto &kv : cf->params) {
    ParamFile &p = kv.second;
    if (p.f && std::fclose(p.f) != 0) rc = fail(BSB200_ERR_BUFFER, "closing %s.bin failed", kv.first.c_str());
    meta += first ? "\n" : ",\n";
    first = false;
    meta += "  \"" + kv.first + "\": {\"file\": \"" + kv.first + ".bin\", \"dtype\": \"" + (p.is_f64 ? "float64" : "int32") +
            "\", \"rows\": " + std::to_string(p.rows) + ", \"cols\": " + std::to_string(p.cols) + ", \"dims\": [";
    for (size_t i = 0; i < p.dims.size(); i++) meta += (i ? ", " : "") + std::to_string(p.dims[i]);
    meta += "]}";
  }
  meta += "\n }\n}\n";
  const std::string path = cf->dir + "/meta.json";
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f || std::fwrite(meta.data(), 1, meta.size(), f) != meta.size()) rc = fail(BSB200_ERR_BUFFER, "cannot write %s", path.c_str());
  if (f) std::fclose(f);
  if (rc == BSB200_OK && !cf->error.empty()) rc = fail(BSB200_ERR_BUFFER, "%s", cf->error.c_str());
  delete cf;
  return rc;
}

}   // extern "C"
