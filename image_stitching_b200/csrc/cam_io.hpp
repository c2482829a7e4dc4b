// cam_io.hpp - text persistence of cameras / indices and the EXIF matrix-string parser.
// Bit-compatible with the reference's serializer (image_stitching/serializer.cpp:7-193).
#pragma once
#include <string>
#include <string_view>
#include <vector>

#include "../../include/image_stitching.h"

namespace isb {

// "[v,v,...]" -> square row-major double matrix of side floor(sqrt(n))   (serializer.cpp:22-36)
bool parse_matrix_str(std::string_view sv, std::vector<double>& out, int& side);
// "[a,b,c;d,e,f;]" writer (serializer.cpp:38-67): default ostream formatting (6 significant digits),
// ',' between columns, ';' after the last column of every row.
std::string serialize_matrix(const double* m, int rows, int cols, bool as_f32);
// reader (serializer.cpp:69-111): strtold tokens, ';' ends a row, ']' ends the matrix; result is float32
bool deserialize_matrix(const std::string& s, std::vector<float>& out, int& rows, int& cols);

bool save_cams(const char* path, const isb_camera* cams, int n);        // serializer.cpp:113-126
bool load_cams(const char* path, std::vector<isb_camera>& cams);        // serializer.cpp:128-167
bool save_indices(const char* path, const int* idx, int n);             // serializer.cpp:169-177
bool load_indices(const char* path, std::vector<int>& idx);             // serializer.cpp:179-193

}  // namespace isb
