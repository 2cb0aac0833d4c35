// boost/filesystem.hpp -- TEST INFRASTRUCTURE.  Stand-in for the four boost::filesystem names irtkReconstructionGPU.cc uses in
// replaceSlices (path, exists, directory_iterator, path::string) over POSIX dirent.  Boost is not in this image; no Boost code here.
#pragma once
#include <dirent.h>
#include <sys/stat.h>
#include <iterator>
#include <ostream>
#include <string>
#include <vector>
#include <algorithm>
#include <memory>

namespace boost { namespace filesystem {
class path {
    std::string s_;
public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    const std::string& string() const { return s_; }
    bool operator<(const path& o) const { return s_ < o.s_; }
    friend std::ostream& operator<<(std::ostream& os, const path& p) { return os << p.s_; }
};
inline bool exists(const path& p) { struct stat st; return ::stat(p.string().c_str(), &st) == 0; }
class directory_iterator {
    std::shared_ptr<std::vector<path>> v_; size_t i_ = 0;
public:
    typedef std::input_iterator_tag iterator_category;
    typedef path value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const path* pointer;
    typedef const path& reference;
    directory_iterator() {}
    explicit directory_iterator(const path& p) : v_(new std::vector<path>())
    {
        if (DIR* d = ::opendir(p.string().c_str())) {
            while (dirent* e = ::readdir(d)) {
                const std::string n = e->d_name;
                if (n != "." && n != "..") v_->push_back(path(p.string() + "/" + n));
            }
            ::closedir(d);
        }
        std::sort(v_->begin(), v_->end());
        if (v_->empty()) v_.reset();
    }
    const path& operator*() const { return (*v_)[i_]; }
    directory_iterator& operator++() { if (v_ && ++i_ >= v_->size()) { v_.reset(); i_ = 0; } return *this; }
    bool operator==(const directory_iterator& o) const { return v_ == o.v_ && i_ == o.i_; }
    bool operator!=(const directory_iterator& o) const { return !(*this == o); }
};
}}
