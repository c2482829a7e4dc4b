#include "cam_io.hpp"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace isb {

bool parse_matrix_str(std::string_view sv, std::vector<double>& out, int& side)
{
    if (sv.size() < 2) return false;
    sv = sv.substr(1, sv.size() - 2);  // strip the enclosing brackets
    std::vector<std::string> items;
    for (size_t pos = sv.find(','); pos != sv.npos; pos = sv.find(',')) {
        items.emplace_back(sv.substr(0, pos));
        sv = sv.substr(pos + 1);
    }
    items.emplace_back(sv);
    side = (int)std::sqrt((double)items.size());
    out.assign((size_t)side * side, 0.0);
    for (int i = 0; i < side * side; ++i) out[i] = std::strtod(items[i].c_str(), nullptr);
    return true;
}

std::string serialize_matrix(const double* m, int rows, int cols, bool as_f32)
{
    std::ostringstream ss;
    ss << "[";
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            if (as_f32) ss << (float)m[r * cols + c];
            else ss << m[r * cols + c];
            ss << (c == cols - 1 ? ";" : ",");
        }
    ss << "]";
    return ss.str();
}

bool deserialize_matrix(const std::string& text, std::vector<float>& out, int& rows, int& cols)
{
    if (text.empty()) return false;
    std::vector<double> values;
    rows = cols = 0;
    const char* p = text.c_str() + 1;  // skip '['
    for (;;) {
        char* end = nullptr;
        values.push_back((double)std::strtold(p, &end));
        if (end == p && *end == '\0') return false;  // malformed: would loop forever
        p = end + 1;
        if (*end == ';') {
            if (rows == 0) ++cols;
            ++rows;
        } else if (rows == 0) {
            ++cols;
        }
        if (*end == '\0') return false;
        if (*p == ']') break;
        if (*p == '\0') return false;
    }
    if (rows <= 0 || cols <= 0 || (size_t)rows * cols > values.size()) return false;
    out.resize((size_t)rows * cols);
    for (int i = 0; i < rows * cols; ++i) out[i] = (float)values[i];
    return true;
}

bool save_cams(const char* path, const isb_camera* cams, int n)
{
    std::ofstream fs(path ? path : "./cams.data");
    if (!fs) return false;
    for (int i = 0; i < n; ++i) {
        const isb_camera& c = cams[i];
        double t[3] = {c.t[0], c.t[1], c.t[2]}, R[9];
        for (int k = 0; k < 9; ++k) R[k] = c.R[k];
        fs << c.aspect << "@" << c.focal << "@" << c.ppx << "@" << c.ppy << "@" << serialize_matrix(t, 3, 1, true) << "@"
           << serialize_matrix(R, 3, 3, true) << std::endl;
    }
    return (bool)fs;
}

bool load_cams(const char* path, std::vector<isb_camera>& cams)
{
    std::ifstream fs(path ? path : "./cams.data");
    if (!fs) return false;
    std::string line;
    while (std::getline(fs, line)) {
        std::string f[6];
        for (int k = 0; k < 5; ++k) {
            size_t pos = line.find('@');
            if (pos == line.npos) return false;
            f[k] = line.substr(0, pos);
            line = line.substr(pos + 1);
        }
        f[5] = line;
        isb_camera c{};
        c.aspect = std::strtod(f[0].c_str(), nullptr);
        c.focal = std::strtod(f[1].c_str(), nullptr);
        c.ppx = std::strtod(f[2].c_str(), nullptr);
        c.ppy = std::strtod(f[3].c_str(), nullptr);
        std::vector<float> m;
        int r, cc;
        if (!deserialize_matrix(f[5], m, r, cc) || r * cc != 9) return false;
        for (int k = 0; k < 9; ++k) c.R[k] = m[k];
        if (!deserialize_matrix(f[4], m, r, cc) || r * cc != 3) return false;
        for (int k = 0; k < 3; ++k) c.t[k] = m[k];
        cams.push_back(c);
    }
    return true;
}

bool save_indices(const char* path, const int* idx, int n)
{
    std::ofstream fs(path ? path : "./indices.data");
    if (!fs) return false;
    for (int i = 0; i < n; ++i) fs << idx[i] << std::endl;
    return (bool)fs;
}

bool load_indices(const char* path, std::vector<int>& idx)
{
    std::ifstream fs(path ? path : "./indices.data");
    if (!fs) return false;
    std::string line;
    while (std::getline(fs, line))
        if (!line.empty()) idx.push_back((int)std::strtol(line.c_str(), nullptr, 10));
    return true;
}

}  // namespace isb
