// CPU check of the SYRK work planner (gpar-at-scale_b200/csrc/syrk_plan.h): every (job, k-block) is
// covered exactly once, slots are consistent, and the cost is balanced over the CTAs.
// usage: syrk_plan_check T NBK with_h num_sms   -> prints "ok <C> <nseg> <max/mean cost> <aligned fraction>"
#include <cstdio>
#include <cstdlib>
#include <map>
#include "../gpar-at-scale_b200/csrc/syrk_plan.h"

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  int T = atoi(argv[1]); long long NBK = atoll(argv[2]); int with_h = atoi(argv[3]); int sms = atoi(argv[4]);
  SyrkPlan pl = plan_syrk(T, NBK, with_h != 0, sms);
  const int J = (int)pl.jobs.size();
  std::vector<std::vector<int>> cover(J, std::vector<int>(NBK, 0));
  if ((int)pl.cta_seg.size() != pl.C + 1 || pl.cta_seg[pl.C] != (int)pl.segs.size() || pl.C > sms || pl.C < 1) { printf("bad cta_seg\n"); return 1; }
  std::vector<double> load(pl.C, 0.0);
  std::vector<int> slot_seen(pl.segs.size(), 0);
  for (int c = 0; c < pl.C; c++) {
    if (pl.cta_seg[c] > pl.cta_seg[c + 1]) { printf("non-monotone cta_seg\n"); return 1; }
    for (int s = pl.cta_seg[c]; s < pl.cta_seg[c + 1]; s++) {
      const SyrkSeg& sg = pl.segs[s];
      if (sg.job < 0 || sg.job >= J || sg.kb0 < 0 || sg.kb1 > NBK || sg.kb0 >= sg.kb1) { printf("bad seg\n"); return 1; }
      const SyrkJob& jb = pl.jobs[sg.job];
      if (jb.a_tile != sg.a_tile || jb.b_tile != sg.b_tile || jb.b_panel != sg.b_panel) { printf("seg/job mismatch\n"); return 1; }
      if (sg.slot < jb.slot0 || sg.slot >= jb.slot0 + jb.nslots || slot_seen[sg.slot]++) { printf("bad slot\n"); return 1; }
      for (int k = sg.kb0; k < sg.kb1; k++) cover[sg.job][k]++;
      bool diag = sg.b_panel == 0 && sg.a_tile == sg.b_tile;
      load[c] += (double)(diag ? SYRK_COST_DIAG : SYRK_COST_REGULAR) * (sg.kb1 - sg.kb0);
    }
  }
  int nslots = 0;
  for (int j = 0; j < J; j++) {
    nslots += pl.jobs[j].nslots;
    for (long long k = 0; k < NBK; k++) if (cover[j][k] != 1) { printf("job %d k-block %lld covered %d times\n", j, k, cover[j][k]); return 1; }
  }
  if (nslots != (int)pl.segs.size()) { printf("slot count mismatch\n"); return 1; }
  double mx = 0, sum = 0;
  for (double l : load) { mx = std::max(mx, l); sum += l; }
  // fraction of the work done in segments whose k range is shared with >= 3 other segments (phase-aligned)
  std::map<std::pair<int, int>, double> ranges;
  for (const SyrkSeg& sg : pl.segs) ranges[{sg.kb0, sg.kb1}] += 1.0;
  double aligned = 0, tot = 0;
  for (const SyrkSeg& sg : pl.segs) { double w = sg.kb1 - sg.kb0; tot += w; if (ranges[{sg.kb0, sg.kb1}] >= 4) aligned += w; }
  printf("ok %d %zu %.4f %.3f\n", pl.C, pl.segs.size(), mx / (sum / sms), aligned / tot);
  return 0;
}
