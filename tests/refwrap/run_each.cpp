// TEST INFRASTRUCTURE. Runs the test functions of one of the reference's tests/cpp/*.cpp files one by one, each in a
// forked child, and prints "name PASS|FAIL". The reference source is compiled UNMODIFIED next to this file with
// -Dmain=reference_main (its own main stops at the first failing assert, and some of its asserts fail against the
// reference library itself); functions are found by their mangled names through dlsym (link with -rdynamic).
//   run_each test_merge_three_step test_merge_with_empty_strings ...
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include <string>

int main(int argc, char** argv) {
    int failures = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string name = argv[i];
        const std::string mangled = "_Z" + std::to_string(name.size()) + name + "v";
        void* fn = dlsym(RTLD_DEFAULT, mangled.c_str());
        if (!fn) {
            printf("%s MISSING\n", name.c_str());
            ++failures;
            continue;
        }
        fflush(stdout);
        const pid_t pid = fork();
        if (pid == 0) {
            // the child's own chatter goes to /dev/null: only the verdict lines matter
            freopen("/dev/null", "w", stdout);
            freopen("/dev/null", "w", stderr);
            reinterpret_cast<void (*)()>(fn)();
            _exit(0);
        }
        int status = 0;
        waitpid(pid, &status, 0);
        const bool ok = WIFEXITED(status) && WEXITSTATUS(status) == 0;
        printf("%s %s\n", name.c_str(), ok ? "PASS" : "FAIL");
        failures += ok ? 0 : 1;
    }
    return failures ? 1 : 0;
}
