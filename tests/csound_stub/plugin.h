// Test infrastructure: a throw-away stand-in for Csound 7 plugin.h (Csound is not installed here), just enough
// surface to syntax-check the reference csound/opcode.cpp against include/ (tests/test_capi.py).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
typedef double MYFLT;
#define OK 0
struct INSDS { int ksmps; };
namespace csnd {
struct Csound {
  void message(const std::string &) {}
  void message(const char *) {}
  int init_error(const std::string &) { return -1; }
  int init_error(const char *) { return -1; }
  int perf_error(const std::string &, void *) { return -1; }
  int perf_error(const char *, void *) { return -1; }
  MYFLT _0dbfs() { return 1.0; }
};
template <typename T> struct Vector {
  T *d; int n;
  int len() { return n; }
  void init(Csound *, int, INSDS *) {}
  T *begin() { return d; }
  T *end() { return d + n; }
};
template <typename T> struct AuxMem {
  std::vector<T> v;
  void allocate(Csound *, int n) { v.resize(n); }
  T *data() { return v.data(); }
  T &operator[](int i) { return v[i]; }
};
struct Table {
  std::vector<MYFLT> v;
  void init(Csound *, MYFLT *) {}
  int len() { return (int)v.size(); }
  MYFLT &operator[](int i) { return v[i]; }
};
struct Args {
  MYFLT vals[16];
  MYFLT &operator[](int i) { return vals[i]; }
  MYFLT *operator()(int i) { return &vals[i]; }
  template <typename T> Vector<T> vector_data(int) { return Vector<T>(); }
};
template <int NOUT, int NIN> struct Plugin {
  Csound *csound; Args inargs, outargs; INSDS *insdshead; int offset, nsmps;
};
struct AudioSig {
  MYFLT *p;
  template <typename P> AudioSig(P *, MYFLT *x) : p(x) {}
  MYFLT &operator[](int i) { return p[i]; }
};
namespace thread { enum { i = 1, k = 2, ik = 3, a = 4, ia = 5 }; }
template <typename T> int plugin(Csound *, const char *, const char *, const char *, int) { return 0; }
}
