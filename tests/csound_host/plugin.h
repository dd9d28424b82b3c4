// Test infrastructure: a minimal FUNCTIONAL stand-in for the Csound 7 plugin framework (plugin.h), enough to host the
// reference's csound/opcode.cpp -- compiled unchanged, from where it lies -- and run its opcodes' init()/aperf()/perf()
// from a C++ driver (tests/cpp/opcode_host_main.cpp). Csound itself is not installed in this image.
// (tests/csound_stub/ is the even smaller set used only for the syntax check in tests/test_capi.py.)
//
// One deliberate extra: Plugin has a member `int i`. Upstream's Cfft::perf / Rfft::perf use an undeclared `i`
// (opcode.cpp:79,87,135,143) and do not compile against the real Csound headers; with this member the whole file
// compiles here, so that the two convolution opcodes -- which are fine upstream -- can be run. The FFT opcodes are
// not exercised by the driver (their perf() is broken upstream whatever `i` resolves to).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

typedef double MYFLT;
#define OK 0
#define NOTOK (-1)
struct INSDS {
  int ksmps;
};

namespace csnd {
struct Csound {
  std::vector<std::string> log;
  std::map<int, std::vector<MYFLT>> tables;  // function tables by number
  MYFLT zerodbfs = 1.0;
  void message(const std::string &s) { log.push_back(s); }
  int init_error(const std::string &s) {
    log.push_back("INIT ERROR: " + s);
    return NOTOK;
  }
  int perf_error(const std::string &s, void *) {
    log.push_back("PERF ERROR: " + s);
    return NOTOK;
  }
  MYFLT _0dbfs() { return zerodbfs; }
};

// k-rate array argument
template <typename T>
struct Vector {
  std::vector<T> *v = nullptr;
  int len() { return v ? (int)v->size() : 0; }
  void init(Csound *, int n, INSDS *) {
    if (v) v->resize(n);
  }
  T *begin() { return v->data(); }
  T *end() { return v->data() + v->size(); }
  T &operator[](int i) { return (*v)[i]; }
};

template <typename T>
struct AuxMem {
  std::vector<T> v;
  void allocate(Csound *, int n) { v.assign(n, T(0)); }
  T *data() { return v.data(); }
  T &operator[](int i) { return v[i]; }
  int len() { return (int)v.size(); }
};

// function table named by an i-time argument holding its number
struct Table {
  std::vector<MYFLT> *t = nullptr;
  void init(Csound *cs, MYFLT *arg) { t = &cs->tables[(int)*arg]; }
  int len() { return t ? (int)t->size() : 0; }
  MYFLT &operator[](int i) { return (*t)[i]; }
};

// opcode arguments: every argument is a pointer to its storage (a scalar for i/k, ksmps values for a-rate,
// a std::vector for k[])
struct Args {
  void *p[16] = {};
  MYFLT &operator[](int i) { return *static_cast<MYFLT *>(p[i]); }
  MYFLT *operator()(int i) { return static_cast<MYFLT *>(p[i]); }
  template <typename T>
  Vector<T> vector_data(int i) {
    Vector<T> r;
    r.v = static_cast<std::vector<T> *>(p[i]);
    return r;
  }
};

template <int NOUT, int NIN>
struct Plugin {
  Csound *csound = nullptr;
  Args outargs, inargs;
  INSDS *insdshead = nullptr;
  uint32_t offset = 0, nsmps = 0;
  int i = 0;  // see the header comment
};

struct AudioSig {
  MYFLT *sig;
  template <typename P>
  AudioSig(P *, MYFLT *s) : sig(s) {}
  MYFLT &operator[](int n) { return sig[n]; }
};

namespace thread {
enum { i = 1, k = 2, ik = 3, a = 4, ia = 5 };
}
template <typename T>
int plugin(Csound *, const char *, const char *, const char *, int) {
  return OK;
}
}  // namespace csnd
