// Concurrent pinned host -> device copy ceiling of the box (torch-free): for N = 1, 2, 4, 8 GPUs, one host thread per GPU
// copies `chunks` x 32 MiB (the chunking of m6a_mil_infer_host_f32) from its own page-locked buffer, all threads started
// together; reports per-GPU and aggregate GB/s.  Variants: buffers first-touched by the copying thread (default) or all by
// the main thread (--main-touch), cudaHostAllocPortable always.  Also D2H with --d2h.
//   nvcc -O2 -std=c++17 -o tools/microbench/h2d_concurrent tools/microbench/h2d_concurrent.cu -lpthread
//   tools/microbench/h2d_concurrent [--json out.json]
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static const size_t kChunk = 32ull << 20;
static const int kChunks = 24;     // 768 MiB per GPU per pass
static const int kPasses = 4;

struct Result { double per_gpu_min, per_gpu_mean, aggregate; };

static Result run(int n_gpus, bool d2h, bool main_touch) {
  std::vector<void*> host(n_gpus, nullptr), dev(n_gpus, nullptr);
  std::vector<cudaStream_t> st(n_gpus);
  for (int g = 0; g < n_gpus; ++g) {
    cudaSetDevice(g);
    cudaMalloc(&dev[g], kChunk * 2);
    cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking);
    if (main_touch) {
      cudaHostAlloc(&host[g], kChunk * kChunks, cudaHostAllocPortable);
      memset(host[g], 1, kChunk * kChunks);
    }
  }
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<double> secs(n_gpus, 0.0);
  std::vector<std::thread> th;
  for (int g = 0; g < n_gpus; ++g)
    th.emplace_back([&, g]() {
      cudaSetDevice(g);
      if (!main_touch) {
        cudaHostAlloc(&host[g], kChunk * kChunks, cudaHostAllocPortable);
        memset(host[g], 1, kChunk * kChunks);
      }
      // warm-up
      cudaMemcpyAsync(dev[g], host[g], kChunk, cudaMemcpyHostToDevice, st[g]);
      cudaStreamSynchronize(st[g]);
      ready.fetch_add(1);
      while (!go.load()) std::this_thread::yield();
      auto t0 = std::chrono::steady_clock::now();
      for (int p = 0; p < kPasses; ++p)
        for (int c = 0; c < kChunks; ++c) {
          char* h = static_cast<char*>(host[g]) + kChunk * c;
          char* d = static_cast<char*>(dev[g]) + kChunk * (c & 1);
          if (d2h) cudaMemcpyAsync(h, d, kChunk, cudaMemcpyDeviceToHost, st[g]);
          else cudaMemcpyAsync(d, h, kChunk, cudaMemcpyHostToDevice, st[g]);
        }
      cudaStreamSynchronize(st[g]);
      secs[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  while (ready.load() < n_gpus) std::this_thread::yield();
  go.store(true);
  for (auto& t : th) t.join();
  const double bytes = double(kChunk) * kChunks * kPasses;
  Result r{1e30, 0.0, 0.0};
  double tmax = 0;
  for (int g = 0; g < n_gpus; ++g) {
    const double gbs = bytes / secs[g] / 1e9;
    r.per_gpu_min = gbs < r.per_gpu_min ? gbs : r.per_gpu_min;
    r.per_gpu_mean += gbs / n_gpus;
    tmax = secs[g] > tmax ? secs[g] : tmax;
  }
  r.aggregate = bytes * n_gpus / tmax / 1e9;
  for (int g = 0; g < n_gpus; ++g) {
    cudaSetDevice(g);
    cudaFreeHost(host[g]);
    cudaFree(dev[g]);
    cudaStreamDestroy(st[g]);
  }
  return r;
}

int main(int argc, char** argv) {
  std::string json_path;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], "--json") && i + 1 < argc) json_path = argv[++i];
  int n_dev = 0;
  cudaGetDeviceCount(&n_dev);
  std::string js = "{\n \"chunk_mib\": 32, \"bytes_per_gpu_per_run\": " + std::to_string(kChunk * kChunks * kPasses) + ",\n";
  const char* names[3] = {"concurrent_h2d_gbs", "concurrent_h2d_gbs_main_thread_touch", "concurrent_d2h_gbs"};
  for (int variant = 0; variant < 3; ++variant) {
    js += std::string(" \"") + names[variant] + "\": {";
    bool first = true;
    for (int n : {1, 2, 4, 8}) {
      if (n > n_dev) break;
      const Result r = run(n, variant == 2, variant == 1);
      printf("%-40s N=%d  per GPU min %.1f mean %.1f GB/s   aggregate %.1f GB/s\n", names[variant], n, r.per_gpu_min, r.per_gpu_mean,
             r.aggregate);
      char buf[256];
      snprintf(buf, sizeof buf, "%s\"%d\": {\"per_gpu\": %.2f, \"per_gpu_mean\": %.2f, \"aggregate\": %.2f}", first ? "" : ", ", n,
               r.per_gpu_min, r.per_gpu_mean, r.aggregate);
      js += buf;
      first = false;
    }
    js += variant < 2 ? "},\n" : "}\n";
  }
  js += "}\n";
  if (!json_path.empty()) {
    FILE* f = fopen(json_path.c_str(), "w");
    if (f) { fputs(js.c_str(), f); fclose(f); }
  }
  return 0;
}
