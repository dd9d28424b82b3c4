// Test infrastructure: runs the reference's Csound opcodes `clconv` (struct Conv) and `cltvconv` (struct TVConv) --
// csound/opcode.cpp compiled UNCHANGED from where it lies (-DOPCODE_CPP=...), hosted by tests/csound_host/plugin.h --
// over raw float32 files, exactly as Csound's performance loop would: init() once, then aperf() once per ksmps block.
// Built twice by __graft_entry__.build(): against include/ + libcl_fft.so (the B200 build) and against the
// reference's own headers + oracle/_ref/libclfft_ref.so (the reference on the host-CPU OpenCL runtime);
// tests/test_opcode_gpu.py runs both on the same inputs and compares the audio they produce.
//
//   opcode_host conv   <parts> <ksmps> <ncycles> <0dbfs> <dir>     reads dir/ir.f32, dir/in.f32          writes dir/out.f32
//   opcode_host tvconv <parts> <ksmps> <ncycles> <0dbfs> <size> <dir>   reads dir/in.f32, dir/in2.f32, dir/frz.f32
// ir.f32: the function table (its length is the IR length: the opcode's optional size / skip arguments are left 0).
// frz.f32: one value per k-cycle, the freeze argument (0 = hold the buffers, opcode.cpp:317-322).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include OPCODE_CPP

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
static void dump_log(csnd::Csound &cs) {
  for (const std::string &s : cs.log) fprintf(stderr, "[csound] %s\n", s.c_str());
}

int main(int argc, char **argv) {
  if (argc < 7) return 2;
  const std::string mode = argv[1];
  const int parts = atoi(argv[2]), ksmps = atoi(argv[3]), ncycles = atoi(argv[4]);
  csnd::Csound cs;
  cs.zerodbfs = atof(argv[5]);
  INSDS ins = {ksmps};
  std::vector<MYFLT> ain(ksmps), ain2(ksmps), aout(ksmps);
  std::vector<float> out((size_t)ksmps * ncycles);
  if (mode == "conv") {
    const std::string dir = argv[6];
    std::vector<float> ir = read_f32(dir + "/ir.f32"), in = read_f32(dir + "/in.f32");
    cs.tables[1].assign(ir.begin(), ir.end());
    csnd::Conv op;
    MYFLT tab = 1, prt = parts, dev = 0, skip = 0, size = 0;
    op.csound = &cs, op.insdshead = &ins;
    op.outargs.p[0] = aout.data();
    op.inargs.p[0] = ain.data(), op.inargs.p[1] = &tab, op.inargs.p[2] = &prt, op.inargs.p[3] = &dev;
    op.inargs.p[4] = &skip, op.inargs.p[5] = &size;
    if (op.init() != OK) {
      dump_log(cs);
      return 1;
    }
    op.offset = 0, op.nsmps = ksmps;
    for (int c = 0; c < ncycles; c++) {
      for (int n = 0; n < ksmps; n++) ain[n] = in[(size_t)c * ksmps + n];
      if (op.aperf() != OK) {
        dump_log(cs);
        return 1;
      }
      for (int n = 0; n < ksmps; n++) out[(size_t)c * ksmps + n] = (float)aout[n];
    }
    op.deinit();
    write_f32(dir + "/out.f32", out);
  } else if (mode == "tvconv" && argc >= 8) {
    const std::string dir = argv[7];
    std::vector<float> in = read_f32(dir + "/in.f32"), in2 = read_f32(dir + "/in2.f32"), frz = read_f32(dir + "/frz.f32");
    csnd::TVConv op;
    MYFLT f1 = 1, f2 = 1, prt = parts, size = atof(argv[6]), dev = 0;
    op.csound = &cs, op.insdshead = &ins;
    op.outargs.p[0] = aout.data();
    op.inargs.p[0] = ain.data(), op.inargs.p[1] = ain2.data(), op.inargs.p[2] = &f1, op.inargs.p[3] = &f2;
    op.inargs.p[4] = &prt, op.inargs.p[5] = &size, op.inargs.p[6] = &dev;
    if (op.init() != OK) {
      dump_log(cs);
      return 1;
    }
    op.offset = 0, op.nsmps = ksmps;
    for (int c = 0; c < ncycles; c++) {
      f1 = f2 = frz[c];
      for (int n = 0; n < ksmps; n++) ain[n] = in[(size_t)c * ksmps + n], ain2[n] = in2[(size_t)c * ksmps + n];
      if (op.aperf() != OK) {
        dump_log(cs);
        return 1;
      }
      for (int n = 0; n < ksmps; n++) out[(size_t)c * ksmps + n] = (float)aout[n];
    }
    op.deinit();
    write_f32(dir + "/out.f32", out);
  } else {
    return 2;
  }
  for (const std::string &s : cs.log)
    if (s.rfind("using device", 0) != 0) printf("%s\n", s.c_str());
  return 0;
}
