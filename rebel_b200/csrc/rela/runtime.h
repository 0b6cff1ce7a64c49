// Actor runtime of the `rela` module: thread loops with pause/resume/terminate and the context that owns their threads.
// Same observable behaviour as the reference's rela::ThreadLoop (rela/thread_loop.h:26-67) and rela::Context
// (rela/context.h:26-85): one std::thread per pushed loop, pause takes effect between units of work, terminated() is true
// once every loop has left mainLoop, the destructor terminates and joins.  (The reference's `started_` flag is never set,
// context.h:28,43 — here it is, so pushing a loop after start() is rejected instead of silently ignored.)
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <string>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace rela {

class ThreadLoop {
 public:
  ThreadLoop() = default;
  ThreadLoop(const ThreadLoop&) = delete;
  ThreadLoop& operator=(const ThreadLoop&) = delete;
  virtual ~ThreadLoop() = default;

  virtual void terminate() {
    terminated_.store(true);
    resume();   // a paused loop must be able to observe termination
  }
  virtual void pause() {
    std::lock_guard<std::mutex> lk(m_);
    paused_ = true;
  }
  virtual void resume() {
    {
      std::lock_guard<std::mutex> lk(m_);
      paused_ = false;
    }
    cv_.notify_all();
  }
  virtual void waitUntilResume() {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [this] { return !paused_; });
  }
  virtual bool terminated() { return terminated_.load(); }
  virtual bool paused() {
    std::lock_guard<std::mutex> lk(m_);
    return paused_;
  }
  virtual void mainLoop() = 0;

 private:
  std::atomic_bool terminated_{false};
  std::mutex m_;
  bool paused_ = false;
  std::condition_variable cv_;
};

class Context {
 public:
  Context() = default;
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  virtual ~Context() {
    for (auto& l : loops_) l->terminate();
    for (auto& t : threads_)
      if (t.joinable()) t.join();
  }

  int pushThreadLoop(std::shared_ptr<ThreadLoop> loop) {
    if (started_) throw std::runtime_error("Context: cannot push a thread loop after start()");
    loops_.push_back(std::move(loop));
    return (int)loops_.size();
  }
  void start() {
    if (started_) return;
    started_ = true;
    for (size_t i = 0; i < loops_.size(); ++i)
      threads_.emplace_back([this, i] {
        // The reference lets an exception escape the worker (-> std::terminate).  Here it is reported and the loop counts
        // as terminated, so the Python side can notice through terminated() / error() instead of losing the process.
        try {
          loops_[i]->mainLoop();
        } catch (const std::exception& e) {
          std::lock_guard<std::mutex> lk(err_m_);
          error_ = e.what();
          std::fprintf(stderr, "[rebel_b200] generator loop %zu stopped: %s\n", i, e.what());
        }
        ++num_terminated_;
      });
  }
  void pause() { for (auto& l : loops_) l->pause(); }
  void resume() { for (auto& l : loops_) l->resume(); }
  void terminate() { for (auto& l : loops_) l->terminate(); }
  bool terminated() { return num_terminated_.load() == (int)loops_.size(); }
  std::string error() {
    std::lock_guard<std::mutex> lk(err_m_);
    return error_;
  }

 private:
  bool started_ = false;
  std::atomic<int> num_terminated_{0};
  std::vector<std::shared_ptr<ThreadLoop>> loops_;
  std::vector<std::thread> threads_;
  std::mutex err_m_;
  std::string error_;
};

}  // namespace rela
