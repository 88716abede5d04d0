/* clock_sampler.c — SM clock / throttle-reason sampling DURING bench.py's timed region, in a native thread.
 *
 * Round 1 polled NVML from a Python thread every millisecond; on the 8-GPU box that poller fought the launch loop for
 * the GIL and one rank lost 6.5 ms inside a 3.8 ms timed region. This sampler runs outside the interpreter: a pthread
 * that calls NVML (dlopen'ed libnvidia-ml.so.1 — no link-time dependency) every `period_us` microseconds between
 * start() and stop(), keeping the samples in a fixed array. Measurement plumbing only; not part of the engine's C-ABI.
 *
 *   int  b200clk_start(const char* gpu_uuid_or_null, int index, int period_us);   0 on success
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
  /* NVML queries take a driver lock the CUDA launch path also wants: a dense poll is affordable for the few
   * milliseconds of the headline region, not for a region of hundreds of milliseconds with thousands of launches (the
   * streaming workload slowed down 1.6x under a constant 250 us poll). So the period backs off: 250 us for the first 16
   * samples, 2 ms up to 64, 20 ms afterwards. */
  int k = 0;
  while (atomic_load(&S.running)) {
    sample_once();
    k++;
    long us = S.period_us;
    if (k > 64) us = us < 20000 ? 20000 : us;
    else if (k > 16) us = us < 2000 ? 2000 : us;
    struct timespec ts = {us / 1000000L, (us % 1000000L) * 1000L};
    nanosleep(&ts, NULL);
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
  S.period_us = period_us > 0 ? period_us : 250;
  sample_once(); /* one sample before the region starts */
  atomic_store(&S.running, 1);
  if (pthread_create(&S.thread, NULL, poll, NULL) != 0) {
    atomic_store(&S.running, 0);
    return -5;
  }
  S.started = 1;
  return 0;
}

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
