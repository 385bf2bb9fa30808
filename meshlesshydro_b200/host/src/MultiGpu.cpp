// see MultiGpu.h
#include "../include/MultiGpu.h"

#include <pthread.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct Control {
    pthread_barrier_t barrier;
    char nccl_id[128];
    volatile int failed;
    volatile pid_t pids[64];
};
int g_planned = 1, g_rank = 0, g_nranks = 1;
bool g_launched = false;
Control *g_ctl = nullptr;
std::vector<pid_t> g_children;
} // namespace

namespace mgpu {

void plan(int nranks) { g_planned = nranks > 1 ? nranks : 1; }
int planned() { return g_planned; }

void *sharedAlloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        std::perror("mgpu::sharedAlloc: mmap");
        std::exit(21);
    }
    return p; // anonymous mappings are zero-filled
}
void sharedFree(void *p, size_t bytes) {
    if (p) munmap(p, bytes ? bytes : 1);
}

void launch() {
    if (g_launched || g_planned <= 1) return;
    g_ctl = (Control *)sharedAlloc(sizeof(Control));
    pthread_barrierattr_t attr;
    pthread_barrierattr_init(&attr);
    pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
    g_nranks = g_planned > 64 ? 64 : g_planned; // the barrier counts the ranks that really exist
    pthread_barrier_init(&g_ctl->barrier, &attr, (unsigned)g_nranks);
    pthread_barrierattr_destroy(&attr);
    g_ctl->pids[0] = getpid();
    std::fflush(stdout);
    std::fflush(stderr);
    for (int r = 1; r < g_nranks; ++r) {
        pid_t pid = fork();
        if (pid < 0) {
            std::perror("mgpu::launch: fork");
            std::exit(21);
        }
        if (pid == 0) { // child = rank r
            g_rank = r;
            g_children.clear();
            g_ctl->pids[r] = getpid();
            break;
        }
        g_ctl->pids[r] = pid;
        g_children.push_back(pid);
    }
    g_launched = true;
}

int rank() { return g_rank; }
int nranks() { return g_nranks; }
bool active() { return g_launched && g_nranks > 1; }

void barrier() {
    if (!active()) return;
    pthread_barrier_wait(&g_ctl->barrier);
    if (g_ctl->failed) {
        std::fprintf(stderr, "rank %d: another rank failed - aborting.\n", g_rank);
        std::_Exit(22);
    }
}

char *ncclId() { return g_ctl ? g_ctl->nccl_id : nullptr; }

// a rank that cannot continue takes the others down with it: they may be blocked in a barrier or an NCCL call
void fail() {
    if (!g_ctl) return;
    g_ctl->failed = 1;
    for (int r = 0; r < g_nranks; ++r)
        if (r != g_rank && g_ctl->pids[r] > 0) kill(g_ctl->pids[r], SIGTERM);
}

int finish(int rc) {
    if (!active()) return rc;
    if (g_rank != 0) {
        std::fflush(stdout);
        std::fflush(stderr);
        std::_Exit(rc);
    }
    int worst = rc;
    for (pid_t pid : g_children) {
        int st = 0;
        if (waitpid(pid, &st, 0) < 0) {
            worst = worst ? worst : 23;
            continue;
        }
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 24;
        if (code != 0 && worst == 0) worst = code;
    }
    return worst;
}

} // namespace mgpu
