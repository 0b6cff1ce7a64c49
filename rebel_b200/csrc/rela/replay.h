// Replay buffer of (value-net query, target values) rows — the consumer side of the data-generation path.
// Python surface identical to the reference's rela::ValuePrioritizedReplay (rela/pybind.cc:126-145,
// rela/prioritized_replay.h:224-506): same constructor, size / num_add / sample / pop_until / load / save / extract / push /
// update_priority, same blocking semantics (producers block while the ring of 1.25 x capacity rows is full, sampling
// evicts the oldest rows down to `capacity`), same on-disk record format (rela/types.cc:87-111).
//
// Not a port: the reference stores one pair of heap-allocated torch tensors per example behind a vector of DataType; at
// GPU generation rates (millions of rows per minute) that is an allocation storm.  Here rows live in two flat float
// arrays (ring buffer), producers append whole waves with one memcpy per block, and a batch is gathered straight into
// a (pinned, when it goes to a GPU) tensor.  Prefetch futures are unnecessary because sampling is a gather of `batch` rows;
// the `prefetch` argument is accepted for API compatibility.
#pragma once
#include <torch/extension.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace rela {

class ValueTransition {
 public:
  ValueTransition() = default;
  ValueTransition(const torch::Tensor& q, const torch::Tensor& v) : query(q), values(v) {}
  torch::Tensor query;
  torch::Tensor values;
};

class ValuePrioritizedReplay {
 public:
  ValuePrioritizedReplay(int capacity, int seed, float alpha, float beta, int prefetch, bool use_priority, bool compressed_values)
      : alpha_(alpha), beta_(beta), prefetch_(prefetch), capacity_(capacity), ring_(int(1.25 * capacity)),
        use_priority_(use_priority), weights_(ring_, 0.f), evicted_(ring_, 0) {
    if (compressed_values) throw std::runtime_error("ValuePrioritizedReplay: compressed_values is not supported by rebel_b200");
    if (capacity <= 0) throw std::runtime_error("ValuePrioritizedReplay: capacity must be positive");
    rng_.seed(seed);
  }

  int size() const { std::lock_guard<std::mutex> lk(m_); return size_; }
  int numAdd() const { return num_add_.load(); }

  // Native producer path: n rows of width q_dim / v_dim; priority may be null (= 1).  Blocks while the ring is full.
  // Returns false if the buffer was closed (shutdown) while waiting.
  bool addRows(const float* q, int q_dim, const float* v, int v_dim, int n, const float* priority) {
    if (n <= 0) return true;
    std::unique_lock<std::mutex> lk(m_);
    ensureWidths(q_dim, v_dim);
    if (n > ring_) throw std::runtime_error("ValuePrioritizedReplay: block larger than the buffer");
    cv_space_.wait(lk, [&] { return size_ + n <= ring_ || closed_; });
    if (closed_) return false;
    double add = 0;
    for (int i = 0; i < n; ++i) {
      const int j = (head_ + size_ + i) % ring_;
      std::copy(q + (size_t)i * q_dim_, q + (size_t)(i + 1) * q_dim_, queries_.begin() + (size_t)j * q_dim_);
      std::copy(v + (size_t)i * v_dim_, v + (size_t)(i + 1) * v_dim_, values_.begin() + (size_t)j * v_dim_);
      float w = priority ? priority[i] : 1.f;
      if (use_priority_) w = std::pow(w, alpha_);
      weights_[j] = w;
      evicted_[j] = 0;
      add += w;
    }
    size_ += n;
    sum_ += add;
    num_add_ += n;
    return true;
  }

  // add(batch, priority) of the reference: slices an [n, ...] transition into rows (prioritized_replay.h:254-261)
  void add(const ValueTransition& batch, const torch::Tensor& priority) {
    auto q = batch.query.to(torch::kCPU, torch::kFloat32).contiguous();
    auto v = batch.values.to(torch::kCPU, torch::kFloat32).contiguous();
    auto p = priority.to(torch::kCPU, torch::kFloat32).contiguous();
    const int n = (int)p.size(0);
    if (q.dim() == 1) q = q.unsqueeze(0);
    if (v.dim() == 1) v = v.unsqueeze(0);
    if (q.size(0) != n || v.size(0) != n) throw std::runtime_error("ValuePrioritizedReplay.add: batch/priority size mismatch");
    addRows(q.data_ptr<float>(), (int)q.size(1), v.data_ptr<float>(), (int)v.size(1), n, p.data_ptr<float>());
  }

  std::tuple<ValueTransition, torch::Tensor> sample(int batchsize, const std::string& device) {
    if (!sampled_ids_.empty() && use_priority_)
      throw std::runtime_error("ValuePrioritizedReplay.sample: previous samples' priority has not been updated");
    std::unique_lock<std::mutex> lk(m_);
    if (size_ <= 0) throw std::runtime_error("ValuePrioritizedReplay.sample: buffer is empty");
    const bool to_gpu = device != "cpu";
    auto opts = torch::TensorOptions().dtype(torch::kFloat32).pinned_memory(to_gpu);
    auto q = torch::empty({batchsize, q_dim_}, opts), v = torch::empty({batchsize, v_dim_}, opts);
    auto w = torch::zeros({batchsize}, torch::kFloat32);
    float* qp = q.data_ptr<float>(); float* vp = v.data_ptr<float>(); float* wp = w.data_ptr<float>();
    std::vector<int> ids(batchsize);
    const int size = size_;
    const double sum = sum_;
    if (!use_priority_) {   // sample_no_priorities_ (prioritized_replay.h:451-486)
      std::uniform_int_distribution<> dist(0, size - 1);
      for (int i = 0; i < batchsize; ++i) {
        const int j = (head_ + dist(rng_)) % ring_;
        ids[i] = j; wp[i] = weights_[j]; evicted_[j] = 0;
        gather(j, qp + (size_t)i * q_dim_, vp + (size_t)i * v_dim_);
      }
    } else {                // stratified proportional sampling (prioritized_replay.h:373-449)
      const float segment = (float)sum / batchsize;
      std::uniform_real_distribution<float> dist(0.0f, segment);
      double acc = 0; int next = 0, id = head_; float wj = 0;
      for (int i = 0; i < batchsize; ++i) {
        float r = std::min((float)sum - 0.1f, dist(rng_) + i * segment);
        while (next < size && !(acc > 0 && acc >= r)) {
          id = (head_ + next) % ring_; wj = weights_[id]; acc += wj; ++next;
        }
        ids[i] = id; wp[i] = wj; evicted_[id] = 0;
        gather(id, qp + (size_t)i * q_dim_, vp + (size_t)i * v_dim_);
      }
    }
    if (size_ > capacity_) popLocked(size_ - capacity_);   // evict oldest down to capacity (prioritized_replay.h:474-477)
    lk.unlock();
    if (use_priority_) {
      sampled_ids_ = ids;
      w = torch::pow(size * (w / (float)sum), -beta_);
      w /= w.max();
    }
    ValueTransition batch(q, v);
    if (to_gpu) {
      auto d = torch::Device(device);
      batch.query = q.to(d, /*non_blocking=*/true);
      batch.values = v.to(d, /*non_blocking=*/true);
      w = w.to(d);
    }
    return std::make_tuple(batch, w);
  }

  void updatePriority(const torch::Tensor& priority) {
    if (priority.size(0) == 0) { sampled_ids_.clear(); return; }
    if ((int)sampled_ids_.size() != priority.size(0)) throw std::runtime_error("update_priority: size mismatch");
    auto p = torch::pow(priority.to(torch::kCPU, torch::kFloat32), alpha_).contiguous();
    const float* pp = p.data_ptr<float>();
    std::lock_guard<std::mutex> lk(m_);
    for (size_t i = 0; i < sampled_ids_.size(); ++i) {
      const int id = sampled_ids_[i];
      if (evicted_[id]) continue;
      sum_ += pp[i] - weights_[id];
      weights_[id] = pp[i];
    }
    sampled_ids_.clear();
  }

  void popUntil(int new_size) {
    std::lock_guard<std::mutex> lk(m_);
    if (size_ > new_size) popLocked(size_ - new_size);
  }

  // Flat binary records: int32 qsize, int32 vsize, float[q], float[v] (rela/types.cc:87-111)
  void save(const std::string& path) {
    std::lock_guard<std::mutex> lk(m_);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    for (int i = 0; i < size_; ++i) {
      const int j = (head_ + i) % ring_;
      std::fwrite(&q_dim_, sizeof(int), 1, f); std::fwrite(&v_dim_, sizeof(int), 1, f);
      std::fwrite(&queries_[(size_t)j * q_dim_], sizeof(float), q_dim_, f);
      std::fwrite(&values_[(size_t)j * v_dim_], sizeof(float), v_dim_, f);
    }
    std::fclose(f);
  }

  void load(const std::string& path, float priority, int max_size, int stride) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<float> q, v;
    for (int added = 0, i = 0;; ++i) {
      if (max_size > 0 && added == max_size) break;
      int qs = 0, vs = 0;
      if (std::fread(&qs, sizeof(int), 1, f) != 1 || std::fread(&vs, sizeof(int), 1, f) != 1) break;
      q.resize(qs); v.resize(vs);
      if ((int)std::fread(q.data(), sizeof(float), qs, f) != qs || (int)std::fread(v.data(), sizeof(float), vs, f) != vs) break;
      if (stride > 1 && i % stride != 0) continue;
      addRows(q.data(), qs, v.data(), vs, 1, &priority);
      ++added;
    }
    std::fclose(f);
  }

  // Whole content as [queries [n,Q], values [n,H], weights [n]] and empty the buffer (prioritized_replay.h:338-345)
  std::vector<torch::Tensor> extract() {
    std::lock_guard<std::mutex> lk(m_);
    const int n = size_;
    auto q = torch::empty({n, std::max(q_dim_, 0)}), v = torch::empty({n, std::max(v_dim_, 0)}), w = torch::empty({n});
    for (int i = 0; i < n; ++i) {
      const int j = (head_ + i) % ring_;
      gather(j, q.data_ptr<float>() + (size_t)i * q_dim_, v.data_ptr<float>() + (size_t)i * v_dim_);
      w.data_ptr<float>()[i] = use_priority_ ? std::pow(weights_[j], 1.f / alpha_) : weights_[j];
    }
    popLocked(n);
    return {q, v, w};
  }

  void push(std::vector<torch::Tensor> data) {
    if (data.size() != 3) throw std::runtime_error("push expects [queries, values, weights]");
    add(ValueTransition(data[0], data[1]), data[2]);
  }

  // Wake producers blocked in addRows for shutdown (not part of the reference surface; used by thread loops).
  void close() {
    { std::lock_guard<std::mutex> lk(m_); closed_ = true; }
    cv_space_.notify_all();
  }

 private:
  void ensureWidths(int q_dim, int v_dim) {
    if (q_dim_ < 0) {
      q_dim_ = q_dim; v_dim_ = v_dim;
      queries_.assign((size_t)ring_ * q_dim_, 0.f);
      values_.assign((size_t)ring_ * v_dim_, 0.f);
    } else if (q_dim != q_dim_ || v_dim != v_dim_) {
      throw std::runtime_error("ValuePrioritizedReplay: row width changed");
    }
  }
  void gather(int j, float* q, float* v) const {
    std::copy(queries_.begin() + (size_t)j * q_dim_, queries_.begin() + (size_t)(j + 1) * q_dim_, q);
    std::copy(values_.begin() + (size_t)j * v_dim_, values_.begin() + (size_t)(j + 1) * v_dim_, v);
  }
  void popLocked(int n) {
    for (int i = 0; i < n; ++i) {
      sum_ -= weights_[head_];
      evicted_[head_] = 1;
      head_ = (head_ + 1) % ring_;
    }
    size_ -= n;
    cv_space_.notify_all();
  }

  const float alpha_, beta_;
  const int prefetch_, capacity_, ring_;
  const bool use_priority_;
  mutable std::mutex m_;
  std::condition_variable cv_space_;
  int q_dim_ = -1, v_dim_ = -1;
  std::vector<float> queries_, values_, weights_;
  std::vector<char> evicted_;
  int head_ = 0, size_ = 0;
  double sum_ = 0;
  bool closed_ = false;
  std::atomic<int> num_add_{0};
  std::vector<int> sampled_ids_;
  std::mt19937 rng_;
};

}  // namespace rela
