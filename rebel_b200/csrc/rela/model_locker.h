// ModelLocker: the hand-over point of value-net weights from the Python trainer to the generator threads.
// Python surface of the reference (rela/pybind.cc:193-195, rela/model_locker.h:54-103): ModelLocker(list_of_models, device)
// and update_model(module).  The reference keeps per-device TorchScript replicas and lets every CFR thread call
// `forward` on them once per iteration; here the value net is evaluated by CUDA kernels inside libcfrb200, so the locker
// snapshots the parameters as ONE flat fp32 buffer (Net2 state_dict order, include/cfrb200.h) with a version counter, and
// each generator loop installs a new version with cfrb_set_weights between two waves.  update_model still refreshes
// the Python replicas (load_state_dict), so code that inspects them keeps working.  One process per GPU: the generator loops of
// the other ranks follow the trainer rank's locker through a stream-ordered ncclBroadcast between two waves (rela_module.cc).
#pragma once
#include <pybind11/pybind11.h>
#include <torch/extension.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace rela {

namespace py = pybind11;

class ModelLocker {
 public:
  ModelLocker(std::vector<py::object> py_models, const std::string& device) : device(device), py_models_(std::move(py_models)) {
    if (py_models_.empty()) throw std::runtime_error("ModelLocker: need at least one model");
    snapshot(py_models_[0]);
    if (this->device.rfind("cuda", 0) != 0 && !std::getenv("CFRB_ACTOR_DEVICE")) {
      // "cpu" lockers (selfplay.cpu_gen_threads) generate on a GPU: the slot is fixed here, once, on the constructing thread
      static std::atomic<int> next{0};
      cpu_slot_ = next++;
      std::fprintf(stderr, "[rebel_b200] ModelLocker(device=\"%s\"): no CPU generation path; this locker generates on a GPU\n",
                   this->device.c_str());
    }
  }

  // Called on the Python thread (GIL held), like the reference (model_locker.h:69-79).
  void updateModel(py::object py_model) {
    py::object sd = py_model.attr("state_dict")();
    for (auto& m : py_models_) m.attr("load_state_dict")(sd);
    snapshot(py_model);
  }

  uint64_t version() const { return version_.load(); }
  std::shared_ptr<const std::vector<float>> weights() const {
    std::lock_guard<std::mutex> lk(m_);
    return weights_;
  }

  // CUDA ordinal the generator loops of this locker run on.  "cuda:i" -> i.  The reference also allows "cpu"
  // (selfplay.cpu_gen_threads); rebel_b200 has no CPU compute path, so CPU lockers are mapped onto the visible GPUs
  // round-robin (CFRB_ACTOR_DEVICE pins one).
  int cudaOrdinal() const {
    if (device.rfind("cuda", 0) == 0) {
      auto pos = device.find(':');
      return pos == std::string::npos ? 0 : std::stoi(device.substr(pos + 1));
    }
    if (const char* e = std::getenv("CFRB_ACTOR_DEVICE")) return std::atoi(e);
    return cpu_slot_ < 0 ? 0 : cpu_slot_;   // taken modulo the device count by the loop
  }

  const std::string device;

 private:
  void snapshot(const py::object& model) {
    static const char* kOrder[] = {"body.0.weight", "body.0.bias", "body.1.weight", "body.1.bias", "body.4.weight", "body.4.bias",
                                   "body.5.weight", "body.5.bias", "output.weight", "output.bias"};
    py::dict sd = model.attr("state_dict")();
    auto flat = std::make_shared<std::vector<float>>();
    // The accelerated net is Net2(n_hidden=256, n_layers=2, use_layer_norm=True) and nothing else: a deeper Net2 (the
    // reference's own default is n_layers=3) also has these ten keys, so every OTHER parameter is rejected instead of being
    // dropped silently (the kernels would evaluate a truncated network), and the hidden width is checked explicitly.
    for (auto item : sd) {
      const std::string key = py::str(item.first);
      bool known = false;
      for (const char* k : kOrder) known |= key == k;
      if (!known)
        throw std::runtime_error("ModelLocker: unexpected parameter '" + key +
                                 "': rebel_b200 accelerates Net2(n_hidden=256, n_layers=2, use_layer_norm=True) only");
    }
    for (const char* key : kOrder) {
      if (!sd.contains(key))
        throw std::runtime_error(std::string("ModelLocker: state_dict has no '") + key +
                                 "' (rebel_b200 accelerates Net2 with n_hidden=256, n_layers=2 and use_layer_norm=True)");
      torch::Tensor t = sd[key].cast<torch::Tensor>().detach().to(torch::kCPU, torch::kFloat32).contiguous();
      if (std::string(key) == "body.0.weight" && (t.dim() != 2 || t.size(0) != 256))
        throw std::runtime_error("ModelLocker: n_hidden must be 256 (got " + std::to_string(t.dim() == 2 ? (long)t.size(0) : -1L) + ")");
      if (std::string(key) == "body.4.weight" && (t.dim() != 2 || t.size(0) != 256 || t.size(1) != 256))
        throw std::runtime_error("ModelLocker: body.4.weight must be [256, 256] (n_hidden=256)");
      const float* p = t.data_ptr<float>();
      flat->insert(flat->end(), p, p + t.numel());
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      weights_ = flat;
    }
    ++version_;
  }

  std::vector<py::object> py_models_;
  mutable std::mutex m_;
  std::shared_ptr<const std::vector<float>> weights_;
  std::atomic<uint64_t> version_{0};
  int cpu_slot_ = -1;
};

}  // namespace rela
