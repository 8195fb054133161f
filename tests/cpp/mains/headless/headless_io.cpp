// TEST BUILD ONLY: file access for the reference's example mains on a machine without OpenCV / jsoncpp.
//  * cv::imread for the raw image containers of the synthetic datasets (see headless_cv.h);
//  * tool::ReadImageSequence / ReadImageSequenceWithPose with the behaviour of the reference's src/Tool/IO.cpp:59-108
//    (associate.txt: "t_rgb rgb t_depth depth" per line; trajectory.txt: one row-major 4x4 per line) -- that translation unit
//    also holds the ScanNet JSON readers and needs jsoncpp, which is why it is not compiled here.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "Tool/IO.h"
namespace cv
{
Mat imread(const std::string &path, int)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { std::cerr << "imread: cannot open " << path << std::endl; return Mat(); }
    char magic[8];
    int32_t hdr[3];
    Mat m;
    if (std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "OPBIMG\0\0", 8) == 0 && std::fread(hdr, 4, 3, f) == 3)
    {
        m.create(hdr[0], hdr[1], hdr[2]);
        const size_t n = (size_t)hdr[0] * hdr[1] * m.elemSize();
        if (std::fread(m.data, 1, n, f) != n) m.release();
    }
    std::fclose(f);
    return m;
}
} // namespace cv
namespace one_piece
{
namespace tool
{
void ReadImageSequence(const std::string &path, std::vector<std::string> &rgb_files, std::vector<std::string> &depth_files)
{
    std::ifstream ifs((path + "/associate.txt").c_str());
    std::string line, t_rgb, t_depth, rgb, depth;
    while (std::getline(ifs, line))
    {
        std::istringstream iss(line);
        iss >> t_rgb >> rgb >> t_depth >> depth;
        rgb_files.push_back(path + "/" + rgb);
        depth_files.push_back(path + "/" + depth);
    }
    std::cout << GREEN << "[ReadImageSequence]::[INFO]::Read " << rgb_files.size() << " images successfully." << RESET << std::endl;
}
void ReadImageSequenceWithPose(const std::string &path, std::vector<std::string> &rgb_files, std::vector<std::string> &depth_files,
                               std::vector<geometry::TransformationMatrix> &poses)
{
    std::ifstream ifs((path + "/trajectory.txt").c_str());
    if (!ifs)
    {
        std::cout << RED << "[ReadImageSequenceWithPose]::[ERROR]::No file named trajectory.txt." << RESET << std::endl;
        return;
    }
    ReadImageSequence(path, rgb_files, depth_files);
    std::string line;
    geometry::TransformationMatrix pose;
    while (std::getline(ifs, line))
    {
        std::istringstream iss(line);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) iss >> pose(r, c);
        poses.push_back(pose);
    }
    if (poses.size() != rgb_files.size())
        std::cout << YELLOW << "[ReadImageSequenceWithPose]::[WARNING]:: The number of images and poses do not match." << RESET << std::endl;
}
} // namespace tool
} // namespace one_piece
