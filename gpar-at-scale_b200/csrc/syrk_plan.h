// syrk_plan.h — host-side work planner of the DMMA panel SYRK (pure C++, unit-tested on the CPU by
// tests/test_syrk_plan.py through tools/syrk_plan_check.cpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

struct SyrkSeg {      // one contiguous k-block range of one tile-job, processed by one CTA
  int a_tile, b_tile;   // M-tile index of the A (rows) and B (cols) operand
  int b_panel;          // 0: B from panel K (G job), 1: B from panel D (H job)
  int kb0, kb1;         // k-block range [kb0, kb1)
  int slot;             // partial-tile slot in the workspace
  int job, pad;
};
struct SyrkJob { int a_tile, b_tile, b_panel, slot0, nslots, pad0, pad1, pad2; };
struct SyrkPlan { std::vector<SyrkJob> jobs; std::vector<SyrkSeg> segs; std::vector<int> cta_seg; int C; };

constexpr int SYRK_COST_REGULAR = 32;   // DMMA per warp and k4-step of a regular 128x128 job
constexpr int SYRK_COST_DIAG = 20;      // ... of a diagonal G job (triangular 16x16-block mapping)

// T M-tiles, NBK k-blocks, with_h: also the full H = K^T D jobs, num_sms CTAs at most.
inline SyrkPlan plan_syrk(int ntiles, int64_t NBK, bool with_h, int num_sms) {
  SyrkPlan pl;
  std::vector<SyrkJob>& jobs = pl.jobs;
  std::vector<SyrkSeg>& segs = pl.segs;
  std::vector<int>& cta_seg = pl.cta_seg;
  typedef SyrkSeg Seg;
  for (int i = 0; i < ntiles; i++)
    for (int j = 0; j <= i; j++) jobs.push_back(SyrkJob{i, j, 0, 0, 0, 0, 0, 0});
  if (with_h)
    for (int i = 0; i < ntiles; i++)
      for (int j = 0; j < ntiles; j++) jobs.push_back(SyrkJob{i, j, 1, 0, 0, 0, 0, 0});
  const int J = (int)jobs.size();
  // Cost-weighted, phase-aligned stream-K.  A k-block of a regular job issues 32 DMMA per warp and
  // k4-step, of a diagonal G job 20.  Every CTA gets the same cost budget T = W / C.  Each job is cut
  // into full pieces of exactly T (piece p of every regular job covers the SAME k range, so the CTAs
  // working on jobs that share an operand tile stream it at the same time and hit in L2); the
  // leftover tails are packed contiguously into the remaining CTAs.
  std::vector<int> cost(J);
  int64_t W = 0;
  for (int j = 0; j < J; j++) {
    cost[j] = (jobs[j].b_panel == 0 && jobs[j].a_tile == jobs[j].b_tile) ? SYRK_COST_DIAG : SYRK_COST_REGULAR;
    W += (int64_t)cost[j] * NBK;
  }
  int C = num_sms;
  if ((int64_t)J * NBK < C) C = (int)((int64_t)J * NBK);
  if (C < 1) C = 1;
  const double T = (double)W / C;
  struct Tail { int job; int64_t kb0; };
  std::vector<Tail> tails;
  int used = 0;
  for (int j = 0; j < J; j++) {
    const double Lj = T / cost[j];                       // piece length in k-blocks
    int f = (int)std::floor((double)NBK / Lj + 1e-9);
    if (used + f > C - 1 && j < J) f = std::max(0, std::min(f, C - 1 - used));   // always leave a CTA for the pool
    int64_t prev = 0;
    for (int q = 0; q < f; q++) {
      int64_t end = (int64_t)std::llround(Lj * (q + 1));
      if (end > NBK) end = NBK;
      if (end > prev) {
        cta_seg.push_back((int)segs.size());
        segs.push_back(Seg{jobs[j].a_tile, jobs[j].b_tile, jobs[j].b_panel, (int)prev, (int)end, 0, j, 0});
        used++;
      }
      prev = end;
    }
    if (prev < NBK) tails.push_back(Tail{j, prev});
  }
  {
    // level 2: the tails of the regular jobs all cover the same k range [K1, NBK); a CTA takes q whole
    // tails plus one partial [K1, K1 + rho), so these CTAs are phase-aligned as well.  What is left
    // (partial remainders, diagonal-job tails) goes to the unaligned pool below.
    std::vector<Tail> reg, rest;
    for (const Tail& t : tails) (cost[t.job] == SYRK_COST_REGULAR ? reg : rest).push_back(t);
    bool same = !reg.empty();
    for (const Tail& t : reg) same = same && t.kb0 == reg[0].kb0;
    if (same && reg.size() >= 2) {
      const int64_t K1 = reg[0].kb0, tau = NBK - K1;
      const double Lk = T / SYRK_COST_REGULAR;
      const int64_t q = (int64_t)std::floor(Lk / (double)tau + 1e-9);
      const int64_t rho = (int64_t)std::floor(Lk - (double)(q * tau) + 1e-9);
      const int64_t per = q + (rho > 0 ? 1 : 0);
      if (q >= 1 && per >= 1) {
        size_t ti = 0;
        while (ti + (size_t)per <= reg.size() && used < C - 1) {
          cta_seg.push_back((int)segs.size());
          for (int64_t i = 0; i < q; i++, ti++) {
            const int j = reg[ti].job;
            segs.push_back(Seg{jobs[j].a_tile, jobs[j].b_tile, jobs[j].b_panel, (int)K1, (int)NBK, 0, j, 0});
          }
          if (rho > 0) {
            const int j = reg[ti].job;
            segs.push_back(Seg{jobs[j].a_tile, jobs[j].b_tile, jobs[j].b_panel, (int)K1, (int)(K1 + rho), 0, j, 0});
            rest.push_back(Tail{j, K1 + rho});
            ti++;
          }
          used++;
        }
        for (; ti < reg.size(); ti++) rest.push_back(reg[ti]);
      } else {
        rest.insert(rest.end(), reg.begin(), reg.end());
      }
    } else {
      rest.insert(rest.end(), reg.begin(), reg.end());
    }
    tails.swap(rest);
    std::sort(tails.begin(), tails.end(), [](const Tail& a, const Tail& b) { return a.job < b.job; });
  }
  {
    // pool: pack the remaining tails into the remaining CTAs, equal cost each
    double pool = 0.0;
    for (const Tail& t : tails) pool += (double)cost[t.job] * (NBK - t.kb0);
    int npool = std::max(1, C - used);
    if (tails.empty()) npool = 0;
    size_t ti = 0; int64_t pos = tails.empty() ? 0 : tails[0].kb0;
    double done = 0.0;
    for (int c = 0; c < npool; c++) {
      const double target = pool * (c + 1) / npool;
      cta_seg.push_back((int)segs.size());
      while (ti < tails.size() && (done < target - 1e-9 || c == npool - 1)) {
        const int j = tails[ti].job;
        int64_t room = (c == npool - 1) ? NBK : pos + (int64_t)std::llround((target - done) / cost[j]);
        int64_t end = std::min<int64_t>(NBK, std::max<int64_t>(room, pos));
        if (end > pos) {
          segs.push_back(Seg{jobs[j].a_tile, jobs[j].b_tile, jobs[j].b_panel, (int)pos, (int)end, 0, j, 0});
          done += (double)cost[j] * (end - pos);
          pos = end;
        }
        if (pos >= NBK) { ti++; if (ti < tails.size()) pos = tails[ti].kb0; }
        else break;
      }
    }
  }
  C = (int)cta_seg.size();
  cta_seg.push_back(0);
  cta_seg[C] = (int)segs.size();
  // slots: segments of a job are consecutive in w-order, so number them in that order
  {
    int slot = 0;
    std::vector<int> order(segs.size());
    for (size_t i = 0; i < segs.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (segs[a].job != segs[b].job) return segs[a].job < segs[b].job;
      return segs[a].kb0 < segs[b].kb0; });
    for (int idx : order) {
      SyrkJob& jb = jobs[segs[idx].job];
      if (jb.nslots == 0) jb.slot0 = slot;
      jb.nslots++;
      segs[idx].slot = slot++;
    }
  }

  pl.C = C;
  return pl;
}
