// ORACLE BUILD SHIM (test infrastructure) -- minimal stand-in for the parts of OpenCV's
// core module that the reference's ORBmatcher.cc and vendored DBoW2 touch.  OpenCV C++
// headers are not installed in this image; this header only exists so that the reference
// sources under /root/reference compile UNMODIFIED into oracle/_ref (see oracle/Makefile).
// Nothing here is product code.
#ifndef ORB_ORACLE_SHIM_CV_CORE_HPP
#define ORB_ORACLE_SHIM_CV_CORE_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv
{
    template <class T>
    struct Point_
    {
        T x, y;
        Point_() : x(0), y(0) {}
        Point_(T a, T b) : x(a), y(b) {}
    };
    typedef Point_<float> Point2f;

    struct KeyPoint
    {
        Point2f pt;
        float size = 0.f, angle = -1.f, response = 0.f;
        int octave = 0, class_id = -1;
    };

    // Row-major dense matrix header with shared storage; row() returns a header aliasing
    // the same buffer like cv::Mat::row.
    class Mat
    {
    public:
        int rows = 0, cols = 0;
        unsigned char *data = nullptr;

        Mat() {}
        Mat(int r, int c, int type) { create(r, c, type); }
        static Mat zeros(int r, int c, int type)
        {
            Mat m(r, c, type);
            if (m.data) std::memset(m.data, 0, (size_t)r * c * m.esz_);
            return m;
        }
        // wraps external memory (no ownership), like cv::Mat(rows, cols, type, void*)
        Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data((unsigned char *)ext), type_(type), esz_(type == CV_32F ? 4 : 1) {}
        void create(int r, int c, int type)
        {
            type_ = type;
            esz_ = (type == CV_32F) ? 4 : 1;
            rows = r;
            cols = c;
            buf_.reset(new unsigned char[(size_t)r * c * esz_ + 1], std::default_delete<unsigned char[]>());
            data = buf_.get();
        }
        void release()
        {
            buf_.reset();
            data = nullptr;
            rows = cols = 0;
        }
        Mat clone() const
        {
            Mat m;
            if (!data) return m;
            m.create(rows, cols, type_);
            std::memcpy(m.data, data, (size_t)rows * cols * esz_);
            return m;
        }
        void copyTo(Mat &dst) const { dst = clone(); }
        Mat row(int i) const
        {
            Mat m;
            m.rows = 1;
            m.cols = cols;
            m.type_ = type_;
            m.esz_ = esz_;
            m.buf_ = buf_;
            m.data = data + (size_t)i * cols * esz_;
            return m;
        }
        bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
        int type() const { return type_; }
        template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * cols * esz_); }
        template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * cols * esz_); }
        template <class T> T &at(int r, int c) { return ptr<T>(r)[c]; }
        template <class T> const T &at(int r, int c) const { return ptr<T>(r)[c]; }

    private:
        int type_ = CV_8U;
        int esz_ = 1;
        std::shared_ptr<unsigned char> buf_;
    };

    // FileStorage / FileNode: only needs to COMPILE (TemplatedVocabulary::save/load are virtual
    // and therefore instantiated); the oracle never calls them.
    class FileNode
    {
    public:
        FileNode operator[](const char *) const { return FileNode(); }
        FileNode operator[](const std::string &) const { return FileNode(); }
        FileNode operator[](int) const { return FileNode(); }
        size_t size() const { return 0; }
        operator int() const { return 0; }
        operator float() const { return 0.f; }
        operator double() const { return 0.0; }
        operator std::string() const { return std::string(); }
    };
    class FileStorage
    {
    public:
        enum { READ = 0, WRITE = 1 };
        FileStorage() {}
        FileStorage(const std::string &, int) {}
        bool isOpened() const { return false; }
        FileNode operator[](const char *) const { return FileNode(); }
        FileNode operator[](const std::string &) const { return FileNode(); }
        void release() {}
    };
    template <class T>
    inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }
} // namespace cv

#endif
