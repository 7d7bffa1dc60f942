// Stand-in for Caffe (not installed; the reference pins no version -- SURVEY.md 8c).  TEST INFRASTRUCTURE, see Eigen/Dense.
// Net<float> is the 3-layer sigmoid encoder of generate_scripts.sh:424-524 with weights read from the raw HF6DW001
// container; its forward pass is delegated to a function the test harness installs (the oracle's fp32 encoder), because
// Caffe's BLAS summation order is unknowable: the encoder is the one stage this build does NOT pin.
#ifndef HF6D_SHIM_CAFFE_HPP
#define HF6D_SHIM_CAFFE_HPP
#include <cstring>
#include <string>
#include <vector>

namespace caffe {

enum Phase { TRAIN = 0, TEST = 1 };

class Caffe {
  public:
    enum Brew { CPU, GPU };
    static void set_mode(Brew) {}
    static void SetDevice(int) {}
};

template <typename T>
class Blob {
  public:
    std::vector<T> d;
    int num_, chan_;
    Blob() : num_(0), chan_(0) {}
    int count() const { return (int)d.size(); }
    T* mutable_cpu_data() { return d.empty() ? 0 : &d[0]; }
    T* mutable_gpu_data() { return mutable_cpu_data(); }
    T data_at(int n, int c, int, int) const { return d[(size_t)n * chan_ + c]; }
};

template <typename T>
class Net {
    Blob<T> in_, out_;
    std::vector<Blob<T>*> in_v_, out_v_;
    std::vector<std::vector<T> > W_, b_;
    std::vector<int> dims_;

  public:
    Net(const std::string& definition_file, Phase phase);
    void CopyTrainedLayersFrom(const std::string& weights_file);
    const std::vector<Blob<T>*>& input_blobs() { return in_v_; }
    const std::vector<Blob<T>*>& ForwardPrefilled();
};

template <typename T>
inline void caffe_copy(int n, const T* src, T* dst) { std::memcpy(dst, src, sizeof(T) * (size_t)n); }

}  // namespace caffe
#endif
