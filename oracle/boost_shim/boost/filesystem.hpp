#pragma once
#include <filesystem>
namespace boost { namespace filesystem {
using std::filesystem::path;
using std::filesystem::exists;
using std::filesystem::is_regular_file;
using std::filesystem::is_directory;
using std::filesystem::remove;
using std::filesystem::remove_all;
using std::filesystem::rename;
using std::filesystem::create_directory;
using std::filesystem::create_directories;
using std::filesystem::current_path;
using std::filesystem::copy_file;
using std::filesystem::file_size;
using std::filesystem::directory_iterator;
} }
