// POSIX-backed stand-in for the few Boost.Filesystem calls in the reference's Utilities.cpp.
#pragma once
#include <string>
#include <ctime>
#include <cstdio>
#include <sys/stat.h>
#include <unistd.h>
#include <dirent.h>
#include <limits.h>
namespace boost { namespace filesystem {
class path {
    std::string s_;
public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    const std::string& string() const { return s_; }
    const char* c_str() const { return s_.c_str(); }
    operator std::string() const { return s_; }
};
typedef path wpath;
inline path current_path() { char b[PATH_MAX]; if (!getcwd(b, sizeof b)) b[0] = 0; return path(b); }
inline void current_path(const path& p) { if (chdir(p.c_str()) != 0) {} }
inline bool exists(const path& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
inline bool is_directory(const path& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool remove(const path& p) { return ::remove(p.c_str()) == 0; }
inline std::time_t last_write_time(const path& p) { struct stat st; if (stat(p.c_str(), &st) != 0) return 0; return st.st_mtime; }
inline bool create_directory(const path& p) { return mkdir(p.c_str(), 0777) == 0; }
inline bool create_directories(const path& p) { return mkdir(p.c_str(), 0777) == 0; }
struct file_status { bool regular; };
inline bool is_regular_file(const file_status& s) { return s.regular; }
class directory_entry { path p_; public: directory_entry() {} directory_entry(const path& p) : p_(p) {} const filesystem::path& path() const { return p_; }
    file_status status() const { struct stat st; file_status f; f.regular = (stat(p_.c_str(), &st) == 0 && S_ISREG(st.st_mode)); return f; } };
class directory_iterator {
    DIR* d_ = nullptr; std::string base_; directory_entry cur_; bool end_ = true;
    void advance() {
        while (d_) { struct dirent* e = readdir(d_); if (!e) { closedir(d_); d_ = nullptr; end_ = true; return; }
            std::string n = e->d_name; if (n == "." || n == "..") continue; cur_ = directory_entry(filesystem::path(base_ + "/" + n)); return; }
    }
public:
    directory_iterator() {}
    explicit directory_iterator(const filesystem::path& p) : base_(p.string()) { d_ = opendir(p.c_str()); end_ = (d_ == nullptr); if (d_) advance(); }
    directory_iterator& operator++() { advance(); return *this; }
    const directory_entry& operator*() const { return cur_; }
    const directory_entry* operator->() const { return &cur_; }
    bool operator!=(const directory_iterator& o) const { return end_ != o.end_; }
    bool operator==(const directory_iterator& o) const { return end_ == o.end_; }
};
inline unsigned long remove_all(const path& p) {
    unsigned long n = 0;
    if (is_directory(p)) { for (directory_iterator it(p); it != directory_iterator(); ++it) n += remove_all(it->path()); rmdir(p.c_str()); return n + 1; }
    return remove(p) ? 1 : 0;
}
} }
