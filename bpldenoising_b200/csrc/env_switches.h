// env_switches.h — the developer switches (BPLTV_*; DESIGN.md "Switches") are read from the environment ONCE, when the
// first context is created (or when bpltv_reload_env() is called — the tests flip them between calls), into a
// process-wide snapshot.  Dispatch code asks the snapshot, never getenv(): no environment scan on the evaluation path.
#pragma once
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

extern char **environ;

namespace bpltv {

struct EnvSnapshot {
    std::mutex mu;
    std::unordered_map<std::string, std::string> vars;
    bool loaded = false;
};
inline EnvSnapshot &env_snapshot()
{
    static EnvSnapshot s;
    return s;
}
inline void env_reload()
{
    EnvSnapshot &s = env_snapshot();
    std::lock_guard<std::mutex> lock(s.mu);
    s.vars.clear();
    for (char **e = environ; e && *e; ++e) {
        if (std::strncmp(*e, "BPLTV_", 6) != 0) continue;
        const char *eq = std::strchr(*e, '=');
        if (eq) s.vars.emplace(std::string(*e, eq - *e), std::string(eq + 1));
    }
    s.loaded = true;
}
// value of a BPLTV_* switch in the snapshot, or nullptr (the pointer stays valid until the next env_reload)
inline const char *env_get(const char *name)
{
    EnvSnapshot &s = env_snapshot();
    if (!s.loaded) env_reload();
    std::lock_guard<std::mutex> lock(s.mu);
    auto it = s.vars.find(name);
    return it == s.vars.end() ? nullptr : it->second.c_str();
}

}  // namespace bpltv
