// batch.cu -- the batch driver of the C ABI (include/b200reg.h, "b200reg_batch_*"): `depth` engine contexts, each on its
// own host thread with its own CUDA streams and memory pool, take alternate jobs.  While one context waits for a PCIe
// upload or polls its LM loop, the kernels of the others keep the SMs busy -- the batch-level concurrency SURVEY.md §8(d)
// names as the lever for the HBM fraction.  Host code only; everything it calls is the public C ABI.
//
// Reference context: the reference registers ONE pair per 2 Hz timer tick, serially, on one thread
// (fast_lio_sam_qn/src/fast_lio_sam_qn.cpp:203-219, LoopClosure is not re-entrant, :81).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200reg.h"

namespace {

struct Job {
  int kind = 0;  // 0 icpAlignment, 1 coarseToFineAlignment
  int count = 0;
  std::vector<const float*> src, tgt;
  std::vector<size_t> src_n, tgt_n;
  size_t stride = 0;
  int on_device = 0;
  b200reg_gicp_params gp;
  b200reg_quatro_params qp;
  b200reg_result* out = nullptr;
  b200reg_quatro_info* qout = nullptr;
  // completion
  int status = 0;
  std::string error;
  bool done = false;
  std::chrono::steady_clock::time_point t_submit, t_done;
};

struct Worker {
  b200reg_ctx* ctx = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<Job>> q;
  bool stop = false;
};

}  // namespace

struct b200reg_batch {
  int device = 0;
  std::vector<std::unique_ptr<Worker>> workers;
  std::mutex mu;  // tickets / completion
  std::condition_variable cv_done;
  std::map<int64_t, std::shared_ptr<Job>> jobs;
  int64_t next_ticket = 0;
  size_t rr = 0;
};

static thread_local std::string g_batch_err;

static void run_worker(b200reg_batch* b, Worker* w) {
  for (;;) {
    std::shared_ptr<Job> j;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->stop || !w->q.empty(); });
      if (w->q.empty()) return;  // stop requested and nothing left
      j = w->q.front();
      w->q.pop_front();
    }
    int rc;
    if (j->kind == 0)
      rc = b200reg_icp_alignment(w->ctx, j->count, j->src.data(), j->src_n.data(), j->tgt.data(), j->tgt_n.data(), j->stride, j->on_device,
                                 &j->gp, j->out);
    else
      rc = b200reg_loop_closure(w->ctx, j->count, j->src.data(), j->src_n.data(), j->tgt.data(), j->tgt_n.data(), j->stride, j->on_device,
                                &j->qp, &j->gp, j->out, j->qout);
    {
      std::lock_guard<std::mutex> lk(b->mu);
      j->status = rc;
      if (rc) j->error = b200reg_last_error();  // this thread's message
      j->t_done = std::chrono::steady_clock::now();
      j->done = true;
    }
    b->cv_done.notify_all();
  }
}

extern "C" {

int b200reg_batch_create(int device, int depth, b200reg_batch** out) {
  if (!out || depth < 1 || depth > 16) return B200REG_EINVAL;
  std::unique_ptr<b200reg_batch> b(new b200reg_batch);
  b->device = device;
  for (int i = 0; i < depth; i++) {
    std::unique_ptr<Worker> w(new Worker);
    const int rc = b200reg_ctx_create(device, &w->ctx);
    if (rc) {  // no GPU: fail loudly, nothing to fall back to
      for (auto& x : b->workers) b200reg_ctx_destroy(x->ctx);
      return rc;
    }
    b->workers.push_back(std::move(w));
  }
  for (auto& w : b->workers) w->th = std::thread(run_worker, b.get(), w.get());
  *out = b.release();
  return B200REG_OK;
}

int b200reg_batch_destroy(b200reg_batch* b) {
  if (!b) return B200REG_OK;
  for (auto& w : b->workers) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
  }
  for (auto& w : b->workers)
    if (w->th.joinable()) w->th.join();
  for (auto& w : b->workers) b200reg_ctx_destroy(w->ctx);
  delete b;
  return B200REG_OK;
}

static int64_t submit(b200reg_batch* b, std::shared_ptr<Job> j) {
  j->t_submit = std::chrono::steady_clock::now();
  int64_t ticket;
  Worker* w;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    ticket = b->next_ticket++;
    b->jobs[ticket] = j;
    w = b->workers[b->rr].get();
    b->rr = (b->rr + 1) % b->workers.size();
  }
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->q.push_back(j);
  }
  w->cv.notify_one();
  return ticket;
}

static std::shared_ptr<Job> make_job(int kind, int count, const float* const* src_xyz, const size_t* src_n, const float* const* tgt_xyz,
                                     const size_t* tgt_n, size_t stride_bytes, int on_device) {
  std::shared_ptr<Job> j(new Job);
  j->kind = kind;
  j->count = count;
  j->src.assign(src_xyz, src_xyz + count);
  j->tgt.assign(tgt_xyz, tgt_xyz + count);
  j->src_n.assign(src_n, src_n + count);
  j->tgt_n.assign(tgt_n, tgt_n + count);
  j->stride = stride_bytes;
  j->on_device = on_device;
  return j;
}

int64_t b200reg_batch_submit_icp(b200reg_batch* b, int count, const float* const* src_xyz, const size_t* src_n, const float* const* tgt_xyz,
                                 const size_t* tgt_n, size_t stride_bytes, int on_device, const b200reg_gicp_params* params,
                                 b200reg_result* out) {
  if (!b || count <= 0 || !src_xyz || !src_n || !tgt_xyz || !tgt_n || !params || !out) return B200REG_EINVAL;
  auto j = make_job(0, count, src_xyz, src_n, tgt_xyz, tgt_n, stride_bytes, on_device);
  j->gp = *params;
  j->out = out;
  return submit(b, j);
}

int64_t b200reg_batch_submit_loop_closure(b200reg_batch* b, int count, const float* const* src_xyz, const size_t* src_n,
                                          const float* const* tgt_xyz, const size_t* tgt_n, size_t stride_bytes, int on_device,
                                          const b200reg_quatro_params* qparams, const b200reg_gicp_params* gparams, b200reg_result* out,
                                          b200reg_quatro_info* quatro_out) {
  if (!b || count <= 0 || !src_xyz || !src_n || !tgt_xyz || !tgt_n || !qparams || !gparams || !out) return B200REG_EINVAL;
  auto j = make_job(1, count, src_xyz, src_n, tgt_xyz, tgt_n, stride_bytes, on_device);
  j->gp = *gparams;
  j->qp = *qparams;
  j->out = out;
  j->qout = quatro_out;
  return submit(b, j);
}

int b200reg_batch_wait(b200reg_batch* b, int64_t ticket, double* latency_ms) {
  if (!b) return B200REG_EINVAL;
  std::shared_ptr<Job> j;
  {
    std::unique_lock<std::mutex> lk(b->mu);
    auto it = b->jobs.find(ticket);
    if (it == b->jobs.end()) return B200REG_EINVAL;  // unknown or already waited for
    j = it->second;
    b->cv_done.wait(lk, [&] { return j->done; });
    b->jobs.erase(it);
  }
  if (latency_ms) *latency_ms = std::chrono::duration<double, std::milli>(j->t_done - j->t_submit).count();
  if (j->status) b200reg_set_last_error(j->error.c_str());
  return j->status;
}

int b200reg_batch_wait_all(b200reg_batch* b) {
  if (!b) return B200REG_EINVAL;
  int first = B200REG_OK;
  for (;;) {
    int64_t t;
    {
      std::lock_guard<std::mutex> lk(b->mu);
      if (b->jobs.empty()) break;
      t = b->jobs.begin()->first;
    }
    const int rc = b200reg_batch_wait(b, t, nullptr);
    if (rc && !first) first = rc;
  }
  return first;
}

int64_t b200reg_batch_launch_count(const b200reg_batch* b) {
  if (!b) return 0;
  int64_t n = 0;
  for (auto& w : b->workers) n += b200reg_ctx_launch_count(w->ctx);
  return n;
}

int b200reg_batch_depth(const b200reg_batch* b) { return b ? (int)b->workers.size() : 0; }

}  // extern "C"
