// simka_count_process_main.cpp -- `simkaCountProcess <simkaCount> args...`: the retry wrapper of the reference
// (src/minikc/SimkaCountProcess.cpp:21-28): runs the command until it exits 0, at most 5 times.
#include <cstdlib>
#include <string>

int main(int argc, char **argv) {
    std::string cmd;
    for (int i = 1; i < argc; i++) { cmd += argv[i]; cmd += " "; }
    if (cmd.empty()) return EXIT_FAILURE;
    for (int attempt = 0; attempt < 5; attempt++)
        if (system(cmd.c_str()) == 0) return EXIT_SUCCESS;
    return EXIT_FAILURE;
}
