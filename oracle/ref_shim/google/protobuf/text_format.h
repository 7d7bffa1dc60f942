// Stand-in for protobuf's TextFormat (libprotobuf C++ is not installed).  TEST INFRASTRUCTURE, see Eigen/Dense.
#ifndef HF6D_SHIM_TEXT_FORMAT_H
#define HF6D_SHIM_TEXT_FORMAT_H
#include <string>
#define GOOGLE_PROTOBUF_VERIFY_VERSION
namespace google {
namespace protobuf {
class TextFormat {
  public:
    template <typename M>
    static bool ParseFromString(const std::string&, M*) { return false; }  // DetectObjects() is compiled, never called
};
}  // namespace protobuf
}  // namespace google
#endif
