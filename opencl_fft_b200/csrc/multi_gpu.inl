// multi_gpu.inl -- one handle, several GPUs (included at the end of capi.cu; declared in include/b200fft.h).
//
// The reference is one object = one channel = one device (cl_conv.cpp:154, cl_fft.cpp:49): nothing is ever exchanged
// between objects, so a set of channels / a batch of transforms shards over the GPUs of a box without any
// communication (no NCCL). A `*_multi` handle owns one ordinary single-device handle per GPU, holding a
// contiguous range of the channels (transforms), and one host worker thread per GPU. A host call hands every worker its
// slice of the caller's buffers; each worker runs the single-device synchronous host entry point (H2D, kernels, D2H on
// that device's own stream), so all devices copy and compute concurrently; the call returns when all have finished.
// Channel g*C/G .. (g+1)*C/G - 1 lives on devices[g] for its whole life: state never moves.
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>

namespace {

struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, stop = false, done = true;
  int rc = B2F_OK;
  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return has_job || stop; });
        if (stop) return;
        std::function<int()> f = std::move(job);
        has_job = false;
        lk.unlock();
        const int r = f();
        lk.lock();
        rc = r;
        done = true;
        cv.notify_all();
      }
    });
  }
  void post(std::function<int()> f) {
    {
      std::lock_guard<std::mutex> lk(m);
      job = std::move(f);
      has_job = true;
      done = false;
    }
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return done; });
    return rc;
  }
  void shutdown() {
    if (!th.joinable()) return;
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv.notify_all();
    th.join();
  }
};

struct Shards {
  int ndev = 0, total = 0;
  std::vector<int> devices;
  std::vector<std::unique_ptr<Worker>> workers;
  // contiguous ranges, the same arithmetic as opencl_fft_b200/shard.py
  int begin(int g, int n) const { return (int)((long long)g * n / ndev); }
  int count(int g, int n) const { return begin(g + 1, n) - begin(g, n); }
  int init(const int *devs, int n, int total_) {
    if (!devs || n < 1 || total_ < n) return B2F_ERR_INVALID_VALUE;
    ndev = n, total = total_;
    devices.assign(devs, devs + n);
    for (int g = 0; g < n; g++)
      for (int k = 0; k < g; k++)
        if (devices[g] == devices[k]) return B2F_ERR_INVALID_VALUE;  // one shard per device
    for (int g = 0; g < n; g++) {
      workers.emplace_back(new Worker);
      workers.back()->start();
    }
    return B2F_OK;
  }
  // run f(g) on every worker, return the first failure
  int fan_out(const std::function<int(int)> &f) {
    for (int g = 0; g < ndev; g++) workers[g]->post([&f, g] { return f(g); });
    int rc = B2F_OK;
    for (int g = 0; g < ndev; g++) {
      const int r = workers[g]->wait();
      if (r && !rc) rc = r;
    }
    return rc;
  }
  void shutdown() {
    for (auto &w : workers) w->shutdown();
    workers.clear();
  }
};

}  // namespace

// ---- partitioned convolution ------------------------------------------------------------------------------------------
struct b2f_pconv_multi {
  Shards sh;
  std::vector<b2f_pconv *> sub;
  int pts = 0, nparts = 0;
};
extern "C" int b2f_pconv_multi_destroy(b2f_pconv_multi *h) {
  if (!h) return B2F_OK;
  h->sh.shutdown();
  for (b2f_pconv *s : h->sub) b2f_pconv_destroy(s);
  delete h;
  return B2F_OK;
}
extern "C" int b2f_pconv_multi_create(b2f_pconv_multi **out, const int *devices, int ndev, int cvs, int pts, int channels) {
  if (!out) return B2F_ERR_INVALID_VALUE;
  *out = nullptr;
  b2f_pconv_multi *h = new (std::nothrow) b2f_pconv_multi;
  if (!h) return B2F_ERR_ALLOC;
  int rc = h->sh.init(devices, ndev, channels);
  h->sub.assign(rc ? 0 : ndev, nullptr);
  h->pts = pts;
  if (!rc) rc = h->sh.fan_out([&](int g) { return b2f_pconv_create(&h->sub[g], h->sh.devices[g], cvs, pts, h->sh.count(g, channels)); });
  if (rc) {
    b2f_pconv_multi_destroy(h);
    return rc;
  }
  h->nparts = b2f_pconv_nparts(h->sub[0]);
  *out = h;
  return B2F_OK;
}
extern "C" int b2f_pconv_multi_nparts(const b2f_pconv_multi *h) { return h ? h->nparts : 0; }
extern "C" int b2f_pconv_multi_reset(b2f_pconv_multi *h) {
  if (!h) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) { return b2f_pconv_reset(h->sub[g]); });
}
extern "C" int b2f_pconv_multi_push_ir_host(b2f_pconv_multi *h, const float *ir, size_t ir_stride) {
  if (!h || !ir) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) {
    return b2f_pconv_push_ir_host(h->sub[g], ir + (size_t)h->sh.begin(g, h->sh.total) * ir_stride, ir_stride);
  });
}
// one device's shard only: `ir` holds the IRs of channels [g*channels/ndev, (g+1)*channels/ndev), `ir_stride` apart
extern "C" int b2f_pconv_multi_push_ir_shard_host(b2f_pconv_multi *h, int g, const float *ir, size_t ir_stride) {
  if (!h || !ir || g < 0 || g >= h->sh.ndev) return B2F_ERR_INVALID_VALUE;
  h->sh.workers[g]->post([&] { return b2f_pconv_push_ir_host(h->sub[g], ir, ir_stride); });
  return h->sh.workers[g]->wait();
}
extern "C" int b2f_pconv_multi_process_host(b2f_pconv_multi *h, float *out, const float *in) {
  if (!h || !out || !in) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) {
    const size_t o = (size_t)h->sh.begin(g, h->sh.total) * h->pts;
    return b2f_pconv_process_host(h->sub[g], out + o, in + o);
  });
}
extern "C" int b2f_pconv_multi_process_tv_host(b2f_pconv_multi *h, float *out, const float *in1, const float *in2) {
  if (!h || !out || !in1 || !in2) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) {
    const size_t o = (size_t)h->sh.begin(g, h->sh.total) * h->pts;
    return b2f_pconv_process_tv_host(h->sub[g], out + o, in1 + o, in2 + o);
  });
}

// ---- direct convolution ------------------------------------------------------------------------------------------------
struct b2f_dconv_multi {
  Shards sh;
  std::vector<b2f_dconv *> sub;
  int vsize = 0;
};
extern "C" int b2f_dconv_multi_destroy(b2f_dconv_multi *h) {
  if (!h) return B2F_OK;
  h->sh.shutdown();
  for (b2f_dconv *s : h->sub) b2f_dconv_destroy(s);
  delete h;
  return B2F_OK;
}
extern "C" int b2f_dconv_multi_create(b2f_dconv_multi **out, const int *devices, int ndev, int irsize, int vsize, int channels,
                                      int max_blocks) {
  if (!out) return B2F_ERR_INVALID_VALUE;
  *out = nullptr;
  b2f_dconv_multi *h = new (std::nothrow) b2f_dconv_multi;
  if (!h) return B2F_ERR_ALLOC;
  int rc = h->sh.init(devices, ndev, channels);
  h->sub.assign(rc ? 0 : ndev, nullptr);
  h->vsize = vsize;
  if (!rc)
    rc = h->sh.fan_out([&](int g) {
      return b2f_dconv_create(&h->sub[g], h->sh.devices[g], irsize, vsize, h->sh.count(g, channels), max_blocks);
    });
  if (rc) {
    b2f_dconv_multi_destroy(h);
    return rc;
  }
  *out = h;
  return B2F_OK;
}
extern "C" int b2f_dconv_multi_reset(b2f_dconv_multi *h) {
  if (!h) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) { return b2f_dconv_reset(h->sub[g]); });
}
extern "C" int b2f_dconv_multi_push_ir_host(b2f_dconv_multi *h, const float *ir, size_t ir_stride) {
  if (!h || !ir) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) {
    return b2f_dconv_push_ir_host(h->sub[g], ir + (size_t)h->sh.begin(g, h->sh.total) * ir_stride, ir_stride);
  });
}
extern "C" int b2f_dconv_multi_process_host(b2f_dconv_multi *h, float *out, const float *in, int nblocks) {
  if (!h || !out || !in || nblocks < 1) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) {
    const size_t o = (size_t)h->sh.begin(g, h->sh.total) * nblocks * h->vsize;
    return b2f_dconv_process_host(h->sub[g], out + o, in + o, nblocks);
  });
}
extern "C" int b2f_dconv_multi_process_tv_host(b2f_dconv_multi *h, float *out, const float *in1, const float *in2) {
  if (!h || !out || !in1 || !in2) return B2F_ERR_INVALID_VALUE;
  return h->sh.fan_out([&](int g) {
    const size_t o = (size_t)h->sh.begin(g, h->sh.total) * h->vsize;
    return b2f_dconv_process_tv_host(h->sub[g], out + o, in1 + o, in2 + o);
  });
}

// ---- FFT batches ---------------------------------------------------------------------------------------------------------
// The transforms of a call are cut into ndev contiguous ranges: transform b of a batch of `batch` runs on device
// g with g*batch/ndev <= b < (g+1)*batch/ndev.
struct b2f_fft_multi {
  Shards sh;
  std::vector<b2f_cfft *> csub;
  std::vector<b2f_rfft *> rsub;
  int N = 0, max_batch = 0;  // complex points per transform
};
static int fft_multi_destroy(b2f_fft_multi *h) {
  if (!h) return B2F_OK;
  h->sh.shutdown();
  for (b2f_cfft *s : h->csub) b2f_cfft_destroy(s);
  for (b2f_rfft *s : h->rsub) b2f_rfft_destroy(s);
  delete h;
  return B2F_OK;
}
static int fft_multi_create(b2f_fft_multi **out, const int *devices, int ndev, int size, int fwd, int max_batch, bool real) {
  if (!out) return B2F_ERR_INVALID_VALUE;
  *out = nullptr;
  b2f_fft_multi *h = new (std::nothrow) b2f_fft_multi;
  if (!h) return B2F_ERR_ALLOC;
  int rc = h->sh.init(devices, ndev, max_batch < ndev ? ndev : max_batch);
  h->N = real ? size / 2 : size;
  h->max_batch = max_batch;
  if (!rc) {
    (real ? (void)h->rsub.assign(ndev, nullptr) : (void)h->csub.assign(ndev, nullptr));
    const int per = (max_batch + ndev - 1) / ndev;
    rc = h->sh.fan_out([&](int g) {
      return real ? b2f_rfft_create(&h->rsub[g], h->sh.devices[g], size, fwd, per)
                  : b2f_cfft_create(&h->csub[g], h->sh.devices[g], size, fwd, per);
    });
  }
  if (rc) {
    fft_multi_destroy(h);
    return rc;
  }
  *out = h;
  return B2F_OK;
}
struct b2f_cfft_multi : b2f_fft_multi {};
struct b2f_rfft_multi : b2f_fft_multi {};
extern "C" int b2f_cfft_multi_create(b2f_cfft_multi **out, const int *devices, int ndev, int N, int fwd, int max_batch) {
  return fft_multi_create(reinterpret_cast<b2f_fft_multi **>(out), devices, ndev, N, fwd, max_batch, false);
}
extern "C" int b2f_cfft_multi_destroy(b2f_cfft_multi *h) { return fft_multi_destroy(h); }
extern "C" int b2f_cfft_multi_exec_host(b2f_cfft_multi *h, float *c, int batch) {
  if (!h || !c || batch < 0) return B2F_ERR_INVALID_VALUE;
  if (batch > h->max_batch) return B2F_ERR_BATCH;
  return h->sh.fan_out([&](int g) {
    const int b0 = h->sh.begin(g, batch), nb = h->sh.count(g, batch);
    return nb ? b2f_cfft_exec_host(h->csub[g], c + (size_t)b0 * 2 * h->N, nb) : B2F_OK;
  });
}
extern "C" int b2f_rfft_multi_create(b2f_rfft_multi **out, const int *devices, int ndev, int size, int fwd, int max_batch) {
  return fft_multi_create(reinterpret_cast<b2f_fft_multi **>(out), devices, ndev, size, fwd, max_batch, true);
}
extern "C" int b2f_rfft_multi_destroy(b2f_rfft_multi *h) { return fft_multi_destroy(h); }
extern "C" int b2f_rfft_multi_exec_host(b2f_rfft_multi *h, float *c, float *r, int batch) {
  if (!h || !c || !r || batch < 0) return B2F_ERR_INVALID_VALUE;
  if (batch > h->max_batch) return B2F_ERR_BATCH;
  return h->sh.fan_out([&](int g) {
    const int b0 = h->sh.begin(g, batch), nb = h->sh.count(g, batch);
    const size_t o = (size_t)b0 * 2 * h->N;  // floats: size reals == size/2 complex per transform
    return nb ? b2f_rfft_exec_host(h->rsub[g], c + o, r + o, nb) : B2F_OK;
  });
}
