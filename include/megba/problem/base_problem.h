// megba/problem/base_problem.h -- BaseProblem: the g2o-style graph container and the only entry point,
// solve() (API of the reference's include/problem/base_problem.h:23-82).
//
// solve() = buildIndex -> algo->solve -> writeBack (src/problem/base_problem.cpp:274-278), implemented on
// the C-ABI of include/megba_b200.h:
//   vertex numbering (setAbsolutePosition, :217-247)   -> dense camera / point indices in ID order
//   edges.forward() once, symbolically                  -> which fused residual kernel to run
//   one ba_solver per ProblemOption::deviceUsed entry   -> one host thread + one NCCL rank per GPU
//   ba_set_params / ba_lm_solve / ba_get_params         -> H2D, LMAlgo::solveCUDA, writeBack (:250-272)
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../megba_b200.h"
#include "../algo/lm_algo.h"
#include "../common.h"
#include "../edge/base_edge.h"
#include "../linear_system/schur_LM_linear_system.h"
#include "../vertex/base_vertex.h"

namespace MegBA {
template <typename T>
class BaseProblem {
  static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "T must be float or double");
  const ProblemOption &problemOption;
  std::unordered_map<int, BaseVertex<T> *> vertices{};
  std::vector<BaseEdge<T> *> edges{};
  const std::unique_ptr<BaseAlgo<T>> algo;
  const std::unique_ptr<BaseLinearSystem<T>> linearSystem;
  std::vector<ba_lm_trace> lastTrace{};

 public:
  explicit BaseProblem(const ProblemOption &option, std::unique_ptr<BaseAlgo<T>> algo_,
                       std::unique_ptr<BaseLinearSystem<T>> linearSystem_)
      : problemOption(option), algo(std::move(algo_)), linearSystem(std::move(linearSystem_)) {
    // same configuration checks as src/problem/base_problem.cpp:60-73
    if (option.N != -1 && option.N != 12)
      throw std::runtime_error("ProblemOption::N must be 12 (9 camera + 3 point parameters) on this path");
    if (option.device != Device::CUDA) throw std::runtime_error("only Device::CUDA is implemented (as in the reference)");
    if (!option.useSchur || option.linearSystemKind != SCHUR || option.solverKind != PCG || option.algoKind != LM)
      throw std::runtime_error("only LM + Schur + PCG is implemented (as in the reference)");
  }

  ~BaseProblem() {
    for (auto &kv : vertices) delete kv.second;
    for (auto *e : edges) delete e;
  }

  const ProblemOption &getProblemOption() const { return problemOption; }
  const std::vector<ba_lm_trace> &getTrace() const { return lastTrace; }

  // takes ownership of `vertex`, like the reference (src/problem/base_problem.cpp:75-80)
  void appendVertex(int ID, BaseVertex<T> *vertex) {
    if (!vertices.insert({ID, vertex}).second) throw std::runtime_error("The ID " + std::to_string(ID) + " exists already.");
  }

  // takes ownership of `edge`
  void appendEdge(BaseEdge<T> &edge) {
    if (edge.size() != 2) throw std::runtime_error("only ONE_CAMERA_ONE_POINT edges are implemented (as in the reference)");
    edges.push_back(&edge);
  }

  BaseVertex<T> &getVertex(int ID) {
    auto it = vertices.find(ID);
    if (it == vertices.end()) throw std::runtime_error("The ID " + std::to_string(ID) + " does not exist in the current graph.");
    return *it->second;
  }
  const BaseVertex<T> &getVertex(int ID) const { return const_cast<BaseProblem *>(this)->getVertex(ID); }

  void eraseVertex(int ID) {
    auto it = vertices.find(ID);
    if (it == vertices.end()) throw std::runtime_error("The ID " + std::to_string(ID) + " does not exist in the current graph.");
    BaseVertex<T> *v = it->second;
    auto dead = std::remove_if(edges.begin(), edges.end(), [&](BaseEdge<T> *e) {
      if (!e->existVertex(*v)) return false;
      delete e;
      return true;
    });
    edges.erase(dead, edges.end());
    delete v;
    vertices.erase(it);
  }

  void solve() {
    if (edges.empty()) throw std::runtime_error("solve(): the graph has no edges");
    // ---- symbolic forward(): the residual of this edge type, either the fused analytical BAL edge (a tag) or a
    // JetVector expression recorded on a tape and compiled to the bytecode the library differentiates on the GPU
    prog::Tape tape;
    prog::Program program;
    bool userProgram = false;
    {
      struct Scope {
        explicit Scope(prog::Tape *t) { prog::current() = t; }
        ~Scope() { prog::current() = nullptr; }
      } scope(&tape);
      edges[0]->bindSymbolic();
      const JVD<T> out = edges[0]->forward();
      if (out.program != ResidualProgram::BAL_ANALYTICAL) {
        if (out.rows() != 2 || out.cols() != 1) throw std::runtime_error("solve(): forward() must return a 2x1 residual");
        program = prog::compile(tape, out(0).id(), out(1).id(), BA_PROGRAM_MAX_SLOTS);
        userProgram = true;
        if (const char *dump = std::getenv("MEGBA_PROGRAM_DUMP")) {  // the compiled residual, for inspection and tests
          if (FILE *f = std::fopen(dump, "w")) {
            std::fprintf(f, "program %d %d %d %d %d\n", program.n_instr(), (int)program.consts.size(), program.n_slots,
                         program.out[0], program.out[1]);
            for (int i = 0; i < program.n_instr(); ++i) {
              const uint8_t *q = &program.code[6 * (size_t)i];
              std::fprintf(f, "%d %d %d %d %d %d\n", q[0], q[1], q[2], q[3], q[4], q[5]);
            }
            for (double k : program.consts) std::fprintf(f, "%.17g\n", k);
            std::fclose(f);
          }
        }
      }
    }
    // ---- vertex numbering: ascending ID inside each kind
    std::vector<std::pair<int, BaseVertex<T> *>> cams, pts;
    for (auto &kv : vertices) {
      if (kv.second->kind() == CAMERA) cams.push_back(kv);
      else if (kv.second->kind() == POINT) pts.push_back(kv);
      else throw std::runtime_error("vertex of unknown kind");
    }
    std::sort(cams.begin(), cams.end());
    std::sort(pts.begin(), pts.end());
    const int64_t nC = (int64_t)cams.size(), nP = (int64_t)pts.size(), nE = (int64_t)edges.size();
    std::vector<T> cam(9 * nC), pt(3 * nP), obs(2 * nE), info(4 * nE);
    std::vector<uint8_t> camFixed(nC, 0), ptFixed(nP, 0);
    bool anyFixed = false;
    for (int64_t c = 0; c < nC; ++c) {
      cams[c].second->absolutePosition = (int)c;
      camFixed[c] = cams[c].second->fixed ? 1 : 0;
      anyFixed = anyFixed || cams[c].second->fixed;
      const auto &e = cams[c].second->getEstimation();
      if (e.size() != 9) throw std::runtime_error("camera estimation must have 9 entries");
      for (int k = 0; k < 9; ++k) cam[9 * c + k] = e(k);
    }
    for (int64_t p = 0; p < nP; ++p) {
      pts[p].second->absolutePosition = (int)p;
      ptFixed[p] = pts[p].second->fixed ? 1 : 0;
      anyFixed = anyFixed || pts[p].second->fixed;
      const auto &e = pts[p].second->getEstimation();
      if (e.size() != 3) throw std::runtime_error("point estimation must have 3 entries");
      for (int k = 0; k < 3; ++k) pt[3 * p + k] = e(k);
    }
    std::vector<int32_t> camIdx(nE), ptIdx(nE);
    bool identityInfo = true;
    for (int64_t j = 0; j < nE; ++j) {
      BaseEdge<T> &e = *edges[j];
      BaseVertex<T> *a = e[0], *b = e[1];
      if (a->kind() != CAMERA) std::swap(a, b);
      if (a->kind() != CAMERA || b->kind() != POINT) throw std::runtime_error("edge must join one camera and one point");
      camIdx[j] = a->absolutePosition;
      ptIdx[j] = b->absolutePosition;
      const auto &m = e._getMeasurement();
      if (m.size() != 2) throw std::runtime_error("measurement must have 2 entries");
      obs[2 * j] = m(0);
      obs[2 * j + 1] = m(1);
      const auto &w = e._getInformation();
      if (w.size() == 4) {
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c) info[4 * j + 2 * r + c] = w(r, c);
        identityInfo = identityInfo && w(0, 0) == T(1) && w(1, 1) == T(1) && w(0, 1) == T(0) && w(1, 0) == T(0);
      } else {
        info[4 * j] = info[4 * j + 3] = T(1);
        info[4 * j + 1] = info[4 * j + 2] = T(0);
      }
    }
    // ---- options
    std::vector<int> devices = problemOption.deviceUsed;
    if (devices.empty()) devices.push_back(0);
    const int ws = (int)devices.size();
    const auto &lmo = algo->algoOption.algoOptionLM;
    const auto &pco = linearSystem->solver->solverOption.solverOptionPCG;
    const ba_lm_options lm{lmo.maxIter, 0, lmo.initialRegion, lmo.epsilon1, lmo.epsilon2};
    const ba_pcg_options pcg{pco.maxIter, 0, pco.tol, pco.refuseRatio};
    const int kind = linearSystem->computeKind() == IMPLICIT ? BA_IMPLICIT : BA_EXPLICIT;
    unsigned char uid[128];
    if (ws > 1 && ba_comm_unique_id(uid) != BA_OK) throw std::runtime_error("ba_comm_unique_id failed");
    std::vector<std::string> errors(ws);
    std::vector<ba_lm_trace> trace((size_t)lmo.maxIter + 2);
    int32_t nTrace = 0;
    // agreement point of the rank threads: returns true when every rank arrived with ok == true
    std::mutex agreeMutex;
    std::condition_variable agreeCv;
    int agreeArrived = 0, agreeGeneration = 0;
    bool agreeAllOk = true, agreeResult = true;
    auto agree = [&](bool ok) {
      if (ws == 1) return ok;
      std::unique_lock<std::mutex> lock(agreeMutex);
      agreeAllOk = agreeAllOk && ok;
      if (++agreeArrived == ws) {
        agreeResult = agreeAllOk;
        agreeArrived = 0;
        agreeAllOk = true;
        ++agreeGeneration;
        agreeCv.notify_all();
        return agreeResult;
      }
      const int gen = agreeGeneration;
      agreeCv.wait(lock, [&] { return agreeGeneration != gen; });
      return agreeResult;
    };
    auto worker = [&](int rank) {
      ba_options opt{};
      opt.dtype = std::is_same<T, double>::value ? BA_F64 : BA_F32;
      opt.compute_kind = kind;
      opt.device = devices[rank];
      opt.rank = rank;
      opt.world_size = ws;
      ba_solver *s = nullptr;
      auto fail = [&](const char *what) { errors[rank] = std::string(what) + ": " + ba_last_error(s); };
      bool ok = ba_create(&opt, nC, nP, nE, camIdx.data(), ptIdx.data(), obs.data(), &s) == BA_OK;
      if (!ok) fail("ba_create");
      // the ranks agree on success before every step that contains a collective: a rank that failed alone would
      // otherwise leave the others waiting in ncclCommInitRank / the first allreduce forever
      if (!agree(ok)) { if (s) ba_destroy(s); return; }
      if (ws > 1 && ba_comm_init(s, uid) != BA_OK) { fail("ba_comm_init"); ok = false; }
      if (ok && !identityInfo && ba_set_information(s, info.data()) != BA_OK) { fail("ba_set_information"); ok = false; }
      if (ok && anyFixed && ba_set_fixed(s, camFixed.data(), ptFixed.data()) != BA_OK) { fail("ba_set_fixed"); ok = false; }
      if (ok && userProgram &&
          ba_set_residual_program(s, program.n_instr(), program.code.data(), (int32_t)program.consts.size(), program.consts.data(),
                                  program.n_slots, program.out) != BA_OK) { fail("ba_set_residual_program"); ok = false; }
      if (ok && ba_set_params(s, cam.data(), pt.data()) != BA_OK) { fail("ba_set_params"); ok = false; }
      if (!agree(ok)) { ba_destroy(s); return; }
      int32_t nt = 0;
      if (ok && ba_lm_solve(s, &lm, &pcg, rank == 0 ? trace.data() : nullptr, (int32_t)trace.size(), &nt, 1) != BA_OK) {
        fail("ba_lm_solve");
        ok = false;
      }
      // every rank takes part in the gather; rank 0's copy is written back
      std::vector<T> camOut(rank == 0 ? 0 : 9 * nC), ptOut(rank == 0 ? 0 : 3 * nP);
      if (ok && ba_get_params(s, rank == 0 ? cam.data() : camOut.data(), rank == 0 ? pt.data() : ptOut.data()) != BA_OK) {
        fail("ba_get_params");
        ok = false;
      }
      if (rank == 0) nTrace = nt;
      ba_destroy(s);
    };
    if (ws == 1) {
      worker(0);
    } else {
      std::vector<std::thread> th;
      for (int r = 0; r < ws; ++r) th.emplace_back(worker, r);
      for (auto &t : th) t.join();
    }
    for (const auto &e : errors)
      if (!e.empty()) throw std::runtime_error(e);
    lastTrace.assign(trace.begin(), trace.begin() + std::min<size_t>(nTrace, trace.size()));
    // ---- writeBack
    for (int64_t c = 0; c < nC; ++c)
      for (int k = 0; k < 9; ++k) cams[c].second->getEstimation()(k) = cam[9 * c + k];
    for (int64_t p = 0; p < nP; ++p)
      for (int k = 0; k < 3; ++k) pts[p].second->getEstimation()(k) = pt[3 * p + k];
  }
};
}  // namespace MegBA
