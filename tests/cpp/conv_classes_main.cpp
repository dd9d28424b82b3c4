// Test infrastructure: drives cl_conv::Clpconv and cl_conv::Cldconv through the C++ class interface (include/cl_conv.h,
// include/cl_dconv.h -- same public interface as the reference's headers, cl_conv.h:156-187, cl_dconv.h:42-65), the way
// a C++ application would: custom error callback, device from clGetDeviceIDs, both convolution() overloads, the
// failed-constructor path. tests/test_dropin_gpu.py runs it on the B200 and compares what it writes with the golden
// vectors produced by the unmodified reference.
//
//   conv_classes pconv <cvs> <pts> <nblocks> <tv 0|1> <dir>    dir/ir.f32 (static), dir/in.f32, dir/in2.f32 (tv) -> dir/out.f32
//   conv_classes dconv <irsize> <vsize> <nblocks> <tv 0|1> <dir>
//   conv_classes failctor
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cl_conv.h"
#include "cl_dconv.h"

struct Sink {
  std::vector<std::string> msgs;
};
static void collect(std::string s, void *d) { static_cast<Sink *>(d)->msgs.push_back(s); }

static std::vector<float> read_f32(const std::string &path) {
  std::vector<float> v;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path.c_str());
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize(n / 4);
  if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2);
  fclose(f);
  return v;
}
static void write_f32(const std::string &path, const std::vector<float> &v) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f || fwrite(v.data(), 4, v.size(), f) != v.size()) exit(2);
  fclose(f);
}
static cl_device_id first_device() {
  cl_device_id ids[32];
  cl_uint num = 0;
  if (clGetDeviceIDs(NULL, CL_DEVICE_TYPE_ALL, 32, ids, &num) != CL_SUCCESS || num == 0) {
    printf("no device\n");
    exit(3);
  }
  return ids[0];
}

template <class Conv>
static int stream(Conv &c, Sink &sink, int block, int nblocks, bool tv, bool has_ir, const std::string &dir) {
  if (c.get_cl_err() != CL_SUCCESS) {
    printf("constructor failed: %s\n", c.cl_error_string(c.get_cl_err()));
    return 1;
  }
  std::vector<float> in = read_f32(dir + "/in.f32"), in2, out((size_t)block * nblocks);
  if (has_ir) {
    std::vector<float> ir = read_f32(dir + "/ir.f32");
    if (c.push_ir(ir.data()) != CL_SUCCESS) return 1;
  }
  if (tv) in2 = read_f32(dir + "/in2.f32");
  for (int b = 0; b < nblocks; b++) {
    const int rc = tv ? c.convolution(&out[(size_t)b * block], &in[(size_t)b * block], &in2[(size_t)b * block])
                      : c.convolution(&out[(size_t)b * block], &in[(size_t)b * block]);
    if (rc != CL_SUCCESS || c.get_cl_err() != CL_SUCCESS) {
      printf("convolution failed at block %d: %s\n", b, c.cl_error_string(rc));
      return 1;
    }
  }
  write_f32(dir + "/out.f32", out);
  printf("ok %d blocks, %zu messages\n", nblocks, sink.msgs.size());
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  Sink sink;
  if (mode == "failctor") {
    // a device that does not exist, and a size the engine rejects: the constructors must not throw or crash,
    // get_cl_err() must be non-zero, the callback must have been told why, and the methods must refuse to run
    cl_conv::Clpconv bad_dev(b2f_cl_device_from_ordinal(99), 4096, 512, collect, &sink);
    cl_conv::Clpconv bad_size(first_device(), 4096, 500, collect, &sink);  // partition size not a power of two
    cl_conv::Cldconv bad_d(b2f_cl_device_from_ordinal(99), 4096, 256, collect, &sink);
    float x[512] = {0}, y[512];
    const int e1 = bad_dev.get_cl_err(), e2 = bad_size.get_cl_err(), e3 = bad_d.get_cl_err();
    const int r1 = bad_dev.convolution(y, x), r2 = bad_size.push_ir(x), r3 = bad_d.convolution(y, x, x);
    printf("errs %d %d %d calls %d %d %d messages %zu\n", e1, e2, e3, r1, r2, r3, sink.msgs.size());
    for (const std::string &m : sink.msgs) printf("msg: %s\n", m.c_str());
    return (e1 > 0 && e2 > 0 && e3 > 0 && r1 > 0 && r2 > 0 && r3 > 0 && sink.msgs.size() >= 3) ? 0 : 1;
  }
  if (argc < 7) return 2;
  const int a = atoi(argv[2]), b = atoi(argv[3]), nblocks = atoi(argv[4]);
  const bool tv = atoi(argv[5]) != 0;
  const std::string dir = argv[6];
  if (mode == "pconv") {
    cl_conv::Clpconv c(first_device(), a, b, collect, &sink);
    return stream(c, sink, b, nblocks, tv, !tv, dir);
  }
  if (mode == "dconv") {
    cl_conv::Cldconv c(first_device(), a, b, collect, &sink);
    return stream(c, sink, b, nblocks, tv, true, dir);
  }
  return 2;
}
