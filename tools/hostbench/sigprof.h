#include <signal.h>
#include <sys/time.h>
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <map>
static uintptr_t g_samples[1 << 20][3]; static volatile int g_ns = 0;
static void on_prof(int, siginfo_t *, void *uc) {
    if (g_ns >= (1 << 20)) return;
    auto &m = ((ucontext_t *)uc)->uc_mcontext;
    uintptr_t rip = m.gregs[REG_RIP], rbp = m.gregs[REG_RBP], rsp = m.gregs[REG_RSP];
    uintptr_t c1 = 0, c2 = 0;
    if (rbp > rsp && rbp < rsp + (1 << 20)) { c1 = ((uintptr_t *)rbp)[1]; uintptr_t r2 = ((uintptr_t *)rbp)[0]; if (r2 > rbp && r2 < rbp + (1 << 20)) c2 = ((uintptr_t *)r2)[1]; }
    int i = g_ns++; g_samples[i][0] = rip; g_samples[i][1] = c1; g_samples[i][2] = c2;
}
static void prof_start() { struct sigaction sa = {}; sa.sa_sigaction = on_prof; sa.sa_flags = SA_SIGINFO | SA_RESTART; sigaction(SIGPROF, &sa, 0); itimerval it = {{0, 1000}, {0, 1000}}; setitimer(ITIMER_PROF, &it, 0); }
static void prof_stop(const char *path) { itimerval it = {}; setitimer(ITIMER_PROF, &it, 0); FILE *f = fopen(path, "w"); std::map<std::pair<uintptr_t, std::pair<uintptr_t, uintptr_t>>, int> h; for (int i = 0; i < g_ns; i++) h[{g_samples[i][0], {g_samples[i][1], g_samples[i][2]}}]++; for (auto &kv : h) fprintf(f, "%lx %d %lx %lx\n", kv.first.first, kv.second, kv.first.second.first, kv.first.second.second); fclose(f); }
