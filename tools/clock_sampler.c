/* clock_sampler.c — SM clock / throttle-reason sampling DURING bench.py's timed region, in a native thread.
 *
 * Round 1 polled NVML from a Python thread every millisecond; on the 8-GPU box that poller fought the launch loop for
 * the GIL and one rank lost 6.5 ms inside a 3.8 ms timed region. This sampler runs outside the interpreter: a pthread
 * that calls NVML (dlopen'ed libnvidia-ml.so.1 — no link-time dependency) every `period_us` microseconds between
 * start() and stop(), keeping the samples in a fixed array. Measurement plumbing only; not part of the engine's C-ABI.
 *
 *   int  b200clk_start(const char* gpu_uuid_or_null, int index, int period_us);   0 on success
 *   void b200clk_arm(void);     the timed region starts now: the in-region schedule counts from here (start() itself
 *                               takes the "before" sample and may be called well ahead of the region)
 *   int  b200clk_stop(unsigned* sm_mhz, unsigned long long* reasons, int capacity, unsigned* sm_max_mhz);  -> #samples
 */
#include <dlfcn.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

typedef void* nvmlDevice_t;
typedef int (*fn_init)(void);
typedef int (*fn_by_uuid)(const char*, nvmlDevice_t*);
typedef int (*fn_by_index)(unsigned, nvmlDevice_t*);
typedef int (*fn_clock)(nvmlDevice_t, int, unsigned*);
typedef int (*fn_reasons)(nvmlDevice_t, unsigned long long*);

#define MAX_SAMPLES 65536
static struct {
  void* lib;
  fn_clock clock_info, max_clock;
  fn_reasons reasons;
  nvmlDevice_t dev;
  pthread_t thread;
  atomic_int running;
  atomic_int armed;
  int started;
  int period_us;
  int n;
  unsigned sm[MAX_SAMPLES];
  unsigned long long mask[MAX_SAMPLES];
} S;

static void sample_once(void) {
  if (S.n >= MAX_SAMPLES) return;
  unsigned mhz = 0;
  unsigned long long r = 0;
  if (S.clock_info(S.dev, 1 /* NVML_CLOCK_SM */, &mhz) != 0) return;
  if (S.reasons) S.reasons(S.dev, &r);
  S.sm[S.n] = mhz;
  S.mask[S.n] = r;
  S.n++;
}

static void* poll(void* arg) {
  (void)arg;
  /* NVML queries take a driver lock that CUDA's launch / copy / NCCL calls also want (measured: with a sample every 250 us the
   * 40 us ncclAllGather that closes bench.py's timed region took 270 us; a streaming workload with thousands of launches ran
   * 1.6x slower). The timed region of the default run is ONE ~1.4 ms kernel with the host merely waiting for it, so the
   * sampler takes its first in-region sample `period_us` (default 400 us) after the start — while that kernel runs and no
   * driver call is in flight — and then backs off: 4 ms, 8 ms, 16 ms, 20 ms ... */
  long us = S.period_us;
  while (atomic_load(&S.running) && !atomic_load(&S.armed)) { /* parked until the region starts */
    struct timespec ts = {0, 20 * 1000L};
    nanosleep(&ts, NULL);
  }
  while (atomic_load(&S.running)) {
    long left = us;
    while (left > 0 && atomic_load(&S.running)) { /* sleep in <= 200 us slices so that stop() returns promptly */
      const long slice = left < 200 ? left : 200;
      struct timespec ts = {0, slice * 1000L};
      nanosleep(&ts, NULL);
      left -= slice;
    }
    if (!atomic_load(&S.running)) break;
    sample_once();
    us = us < 4000 ? 4000 : (us * 2 > 20000 ? 20000 : us * 2);
  }
  return NULL;
}

int b200clk_start(const char* uuid, int index, int period_us) {
  if (S.started) return -1;
  if (!S.lib) {
    S.lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
    if (!S.lib) return -2;
    fn_init init = (fn_init)dlsym(S.lib, "nvmlInit_v2");
    fn_by_uuid by_uuid = (fn_by_uuid)dlsym(S.lib, "nvmlDeviceGetHandleByUUID");
    fn_by_index by_index = (fn_by_index)dlsym(S.lib, "nvmlDeviceGetHandleByIndex_v2");
    S.clock_info = (fn_clock)dlsym(S.lib, "nvmlDeviceGetClockInfo");
    S.max_clock = (fn_clock)dlsym(S.lib, "nvmlDeviceGetMaxClockInfo");
    S.reasons = (fn_reasons)dlsym(S.lib, "nvmlDeviceGetCurrentClocksEventReasons");
    if (!S.reasons) S.reasons = (fn_reasons)dlsym(S.lib, "nvmlDeviceGetCurrentClocksThrottleReasons");
    if (!init || !by_index || !S.clock_info || init() != 0) return -3;
    int ok = -1;
    if (uuid && by_uuid) ok = by_uuid(uuid, &S.dev);
    if (ok != 0) ok = by_index((unsigned)index, &S.dev);
    if (ok != 0) return -4;
  }
  S.n = 0;
  S.period_us = period_us > 0 ? period_us : 400;
  sample_once(); /* one sample before the region starts */
  atomic_store(&S.armed, 0);
  atomic_store(&S.running, 1);
  if (pthread_create(&S.thread, NULL, poll, NULL) != 0) {
    atomic_store(&S.running, 0);
    return -5;
  }
  S.started = 1;
  return 0;
}

void b200clk_arm(void) { atomic_store(&S.armed, 1); }

int b200clk_stop(unsigned* sm_mhz, unsigned long long* reasons, int capacity, unsigned* sm_max_mhz) {
  if (!S.started) return 0;
  atomic_store(&S.running, 0);
  pthread_join(S.thread, NULL);
  S.started = 0;
  sample_once(); /* and one after it ended */
  int n = S.n < capacity ? S.n : capacity;
  for (int i = 0; i < n; i++) {
    sm_mhz[i] = S.sm[i];
    reasons[i] = S.mask[i];
  }
  if (sm_max_mhz) {
    unsigned mx = 0;
    if (S.max_clock) S.max_clock(S.dev, 1, &mx);
    *sm_max_mhz = mx;
  }
  return n;
}
