/* Debug helper: gcc -shared -fPIC -o /tmp/libsegvbt.so tools/segv_backtrace.c ; load it (ctypes.CDLL) to get a NATIVE
 * backtrace on SIGSEGV (faulthandler only shows Python frames). */
#include <execinfo.h>
#include <signal.h>
#include <string.h>
#include <unistd.h>
static void handler(int sig) {
    void *frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "\n==== native backtrace (SIGSEGV) ====\n";
    write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
__attribute__((constructor)) static void install(void) {
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_handler = handler;
    sigaction(SIGSEGV, &sa, 0);
}
