// TEST INFRASTRUCTURE (oracle/_ref build only).  A minimal stand-in for the OpenCV 3 C++ API, just large enough to
// compile the reference's vendored 3rdparty/line_descriptor sources UNMODIFIED, where they lie under /root/reference,
// into oracle/_ref/ (see oracle/ref_build/Makefile).  The image has no OpenCV C++ headers.  Only what the LBD / KeyLine
// path executes is functional; the image-processing primitives that path calls (GaussianBlur 5x5, Sobel 3x3,
// createLineSegmentDetector, LineIterator::count) are supplied by the cv2-pinned C restatements of oracle/*.c.
// Everything else (EDLines, matcher, drawing) only has to compile and aborts if it is ever called.
#ifndef PLF_OPENCV_STUB_CORE_HPP
#define PLF_OPENCV_STUB_CORE_HPP
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_EXPORTS_W_SIMPLE
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_PROP_RW
#define CV_PROP
#define CV_WRAP_AS(x)

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;
typedef int64_t int64;
typedef uint64_t uint64;

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {
using std::abs;
using std::max;
using std::min;
using std::sqrt;
typedef std::string String;

[[noreturn]] inline void plf_stub_abort(const char* what) {
  fprintf(stderr, "opencv stub: %s is not implemented (not on the LBD / KeyLine path)\n", what);
  abort();
}

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
namespace Error { enum Code { StsOk = 0, StsError = -2, StsBadArg = -5, StsBadSize = -201, StsAssert = -215 }; }
inline void error(int, const String& msg, const char* func, const char* file, int line) {
  std::ostringstream o; o << file << ":" << line << " " << func << ": " << msg; throw Exception(o.str());
}
inline String format(const char* fmt, ...) { return String(fmt); }
#define CV_Error(code, msg) cv::error(code, msg, __func__, __FILE__, __LINE__)
#define CV_Assert(expr) do { if (!(expr)) cv::error(cv::Error::StsAssert, #expr, __func__, __FILE__, __LINE__); } while (0)
#define CV_DbgAssert(expr)

template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline short saturate_cast<short>(double v) { int i = cvRound(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
  Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
  bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
  T area() const { return width * height; }
};
typedef Size_<int> Size;
typedef Size_<float> Size2f;

struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  static Scalar all(double v) { return Scalar(v, v, v, v); }
  double operator[](int i) const { return val[i]; }
};
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Range { int start, end; Range() : start(0), end(0) {} Range(int s, int e) : start(s), end(e) {} static Range all() { return Range(INT32_MIN, INT32_MAX); } };

template <typename T, int n> struct Vec {
  T val[n];
  Vec() { for (int i = 0; i < n; ++i) val[i] = T(); }
  Vec(T a, T b) { val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
  Vec(T a, T b, T c, T d) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;
typedef Vec<float, 2> Vec2f;
typedef Vec<uchar, 3> Vec3b;

template <typename T> struct DataType { enum { type = -1 }; };
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<schar> { enum { type = CV_8SC1 }; };
template <> struct DataType<ushort> { enum { type = CV_16UC1 }; };
template <> struct DataType<short> { enum { type = CV_16SC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };

inline int plf_elem_size(int type) {
  static const int d[8] = {1, 1, 2, 2, 4, 4, 8, 0};
  return d[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
}

class Mat {
 public:
  int flags, dims, rows, cols;
  uchar* data;
  struct Step { size_t p; Step() : p(0) {} operator size_t() const { return p; } Step& operator=(size_t v) { p = v; return *this; } size_t operator[](int i) const { return i == 0 ? p : 0; } } step;
  std::shared_ptr<std::vector<uchar>> buf;

  Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& v) : Mat() { create(r, c, type); setTo(v); }
  Mat(Size s, int type, const Scalar& v) : Mat() { create(s.height, s.width, type); setTo(v); }
  Mat(int r, int c, int type, void* d, size_t st = 0) : flags(type), dims(2), rows(r), cols(c), data((uchar*)d) { step = st ? st : (size_t)c * plf_elem_size(type); }
  Mat(const Mat& m, const Rect&) : Mat() { (void)m; plf_stub_abort("Mat(Mat, Rect)"); }
  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && this->type() == type && buf) return;
    flags = type; rows = r; cols = c; step = (size_t)c * plf_elem_size(type);
    buf = std::make_shared<std::vector<uchar>>((size_t)r * step.p + 64, (uchar)0);
    data = buf->data();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  void release() { buf.reset(); data = nullptr; rows = cols = 0; }
  int type() const { return flags & 4095; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize() const { return (size_t)plf_elem_size(type()); }
  size_t elemSize1() const { return elemSize() / channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step.p == (size_t)cols * elemSize(); }
  Size size() const { return Size(cols, rows); }
  size_t total() const { return (size_t)rows * cols; }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& m) const {
    if (empty()) { m.release(); return; }
    m.create(rows, cols, type());
    for (int r = 0; r < rows; ++r) memcpy(m.data + (size_t)r * m.step.p, data + (size_t)r * step.p, (size_t)cols * elemSize());
  }
  void convertTo(Mat&, int, double = 1, double = 0) const { plf_stub_abort("Mat::convertTo"); }
  Mat& setTo(const Scalar& v) {
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols * channels(); ++c) {
        uchar* p = data + (size_t)r * step.p + (size_t)c * elemSize1();
        switch (depth()) {
          case CV_8U: *p = (uchar)v.val[0]; break;
          case CV_8S: *(schar*)p = (schar)v.val[0]; break;
          case CV_16U: *(ushort*)p = (ushort)v.val[0]; break;
          case CV_16S: *(short*)p = (short)v.val[0]; break;
          case CV_32S: *(int*)p = (int)v.val[0]; break;
          case CV_32F: *(float*)p = (float)v.val[0]; break;
          default: *(double*)p = v.val[0];
        }
      }
    return *this;
  }
  Mat& operator=(const Scalar& v) { return setTo(v); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step.p; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step.p; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.p); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.p); }
  template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.p))[c]; }
  template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.p))[c]; }
  template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step.p); }
  template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step.p); }
  template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
  template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
  Mat row(int r) const { Mat m; m.flags = flags; m.dims = 2; m.rows = 1; m.cols = cols; m.data = data + (size_t)r * step.p; m.step = step.p; m.buf = buf; return m; }
  Mat col(int) const { plf_stub_abort("Mat::col"); }
  Mat t() const { plf_stub_abort("Mat::t"); }
  Mat inv(int = 0) const { plf_stub_abort("Mat::inv"); }
  Mat operator()(const Rect& r) const { return Mat(*this, r); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0)); }
  static Mat zeros(Size s, int type) { return Mat(s, type, Scalar(0)); }
  static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
  static Mat ones(Size s, int type) { return Mat(s, type, Scalar(1)); }
};

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() : Mat() {}
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(const Mat& m) : Mat(m) {}
  Mat_& operator=(const Mat& m) { Mat::operator=(m); return *this; }
  T& operator()(int r, int c) { return this->template at<T>(r, c); }
  const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
  T* operator[](int r) { return this->template ptr<T>(r); }
  const T* operator[](int r) const { return this->template ptr<T>(r); }
};

// ---- array proxies -------------------------------------------------------------------------------------------------
class _InputArray {
 public:
  const Mat* m;
  _InputArray() : m(nullptr) {}
  _InputArray(const Mat& mm) : m(&mm) {}
  template <typename T> _InputArray(const std::vector<T>&) : m(nullptr) {}
  Mat getMat(int = -1) const { return m ? *m : Mat(); }
  bool empty() const { return !m || m->empty(); }
  int type() const { return m ? m->type() : 0; }
  Size size() const { return m ? m->size() : Size(); }
};
class _OutputArray {
 public:
  Mat* m;
  std::vector<Vec4f>* v4f;
  _OutputArray() : m(nullptr), v4f(nullptr) {}
  _OutputArray(Mat& mm) : m(&mm), v4f(nullptr) {}
  _OutputArray(std::vector<Vec4f>& v) : m(nullptr), v4f(&v) {}
  template <typename T> _OutputArray(std::vector<T>&) : m(nullptr), v4f(nullptr) {}
  bool needed() const { return m != nullptr || v4f != nullptr; }
  void create(int r, int c, int type) const { if (m) m->create(r, c, type); }
  void create(Size s, int type) const { if (m) m->create(s, type); }
  Mat getMat(int = -1) const { return m ? *m : Mat(); }
  Mat& getMatRef() const { return *m; }
  operator Mat&() const { return *m; }
  void release() const { if (m) m->release(); }
};
typedef const _InputArray& InputArray;
typedef InputArray InputArrayOfArrays;
typedef const _OutputArray& OutputArray;
typedef OutputArray OutputArrayOfArrays;
typedef const _OutputArray& InputOutputArray;
inline const _OutputArray& noArray() { static _OutputArray a; return a; }

// ---- smart pointer / Algorithm / persistence ------------------------------------------------------------------------
template <typename T> class Ptr : public std::shared_ptr<T> {
 public:
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  template <typename U> Ptr(const std::shared_ptr<U>& o) : std::shared_ptr<T>(o) {}
  template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
  bool empty() const { return !this->get(); }
  void release() { this->reset(); }
  operator T*() const { return this->get(); }
};
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

class FileNode {
 public:
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](const String&) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  bool empty() const { return true; }
  bool isNone() const { return true; }
  operator int() const { return 0; }
  operator float() const { return 0.f; }
  operator double() const { return 0.0; }
  operator std::string() const { return std::string(); }
};
template <typename T> inline void operator>>(const FileNode&, T&) {}
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const String&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode root() const { return FileNode(); }
  FileNode getFirstTopLevelNode() const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](const String&) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

class Algorithm {
 public:
  virtual ~Algorithm() {}
  virtual void clear() {}
  virtual void write(FileStorage&) const {}
  virtual void read(const FileNode&) {}
  virtual bool empty() const { return false; }
  virtual void save(const String&) const {}
  virtual String getDefaultName() const { return String("my_object"); }
};

struct KeyPoint {
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
  int queryIdx, trainIdx, imgIdx; float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  bool operator<(const DMatch& m) const { return distance < m.distance; }
};

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };

inline int64 getTickCount() { return 0; }
inline double getTickFrequency() { return 1.0; }

// arithmetic used only by code that is off the LBD path: compile-only
inline Mat abs(const Mat&) { plf_stub_abort("cv::abs(Mat)"); }
inline Mat operator+(const Mat&, const Mat&) { plf_stub_abort("Mat + Mat"); }
inline Mat operator-(const Mat&, const Mat&) { plf_stub_abort("Mat - Mat"); }
inline Mat operator*(const Mat&, double) { plf_stub_abort("Mat * s"); }
inline Mat operator*(const Mat&, const Mat&) { plf_stub_abort("Mat * Mat"); }
inline Mat operator*(double, const Mat&) { plf_stub_abort("s * Mat"); }
inline Mat operator/(const Mat&, double) { plf_stub_abort("Mat / s"); }
inline void add(InputArray, InputArray, OutputArray, InputArray = _InputArray(), int = -1) { plf_stub_abort("cv::add"); }
inline void compare(InputArray, InputArray, OutputArray, int) { plf_stub_abort("cv::compare"); }
inline double norm(InputArray, int = NORM_L2, InputArray = _InputArray()) { plf_stub_abort("cv::norm"); }
inline double norm(InputArray a_, InputArray b_, int type = NORM_L2, InputArray = _InputArray()) {
  Mat a = a_.getMat(), b = b_.getMat();   // functional for what src/mapFeatures.cpp:63,133 calls: NORM_HAMMING on CV_8U rows
  if (type != NORM_HAMMING || a.depth() != CV_8U || a.type() != b.type() || a.rows != b.rows || a.cols != b.cols) plf_stub_abort("cv::norm (only NORM_HAMMING on CV_8U)");
  int d = 0;
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols * a.channels(); ++c) d += __builtin_popcount((unsigned)(a.ptr(r)[c] ^ b.ptr(r)[c]));
  return (double)d;
}
inline int countNonZero(InputArray) { plf_stub_abort("cv::countNonZero"); }
}  // namespace cv
#endif
