#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#define MAXS 400000
#define DEPTH 24
static void* samples[MAXS][DEPTH];
static int depth[MAXS];
static volatile int ns = 0;
static void handler(int sig) {
  if (ns >= MAXS) return;
  int i = ns++;
  depth[i] = backtrace(samples[i], DEPTH);
}
__attribute__((constructor)) static void init(void) {
  void* tmp[4]; backtrace(tmp, 4);   /* load libgcc now, not in the handler */
  struct sigaction sa; memset(&sa, 0, sizeof(sa)); sa.sa_handler = handler; sa.sa_flags = SA_RESTART;
  sigaction(SIGPROF, &sa, 0);
  struct itimerval it; it.it_interval.tv_sec = 0; it.it_interval.tv_usec = 10000; it.it_value = it.it_interval;
  setitimer(ITIMER_PROF, &it, 0);
}
__attribute__((destructor)) static void fini(void) {
  struct itimerval it; memset(&it, 0, sizeof(it)); setitimer(ITIMER_PROF, &it, 0);
  const char* path = getenv("SPROF_OUT"); if (!path) path = "sprof.out";
  FILE* f = fopen(path, "w"); if (!f) return;
  FILE* m = fopen("/proc/self/maps", "r"); char line[512];
  while (m && fgets(line, sizeof(line), m)) fprintf(f, "MAP %s", line);
  if (m) fclose(m);
  for (int i = 0; i < ns; ++i) { fprintf(f, "S"); for (int d = 2; d < depth[i]; ++d) fprintf(f, " %p", samples[i][d]); fprintf(f, "\n"); }
  fclose(f);
}
