// engine.h -- host-side plan of the SG-GAN step: layer geometry, frame layouts, workspace carving
// and the launch sequences for generator / discriminator forward, both backward passes and Adam.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/sggan.h"
#include "conv_gemm_tc.h"
#include "glue.h"

namespace sggan {

enum LayerType { LT_S1 = 0, LT_S2 = 1, LT_DECONV = 2, LT_WIN_C1 = 3, LT_OUT7 = 4, LT_WIN_H0 = 5 };
enum PadMode { PAD_VALID = 0, PAD_ZERO = 1, PAD_REFLECT = 2 };

struct TensorInfo {
  int64_t offset, numel;
  int rank;
  int64_t shape[4];
};

struct Layer {
  LayerType type;
  int k;
  PadMode pad;
  int Cin, Cout;  // logical channels
  int Hin, Win, Hout, Wout;
  bool has_norm;
  int act;  // activation after the norm (glue) or in the conv epilogue (no norm)
  float alpha;
  int nb, nbv;                // forward batch, backward (virtual) batch
  int ti_w, ti_b, ti_g, ti_be;  // tensor indices inside the net (-1 if absent)
  int P;                      // shared pitch of the input frame and the output-gradient frame
  float* nr_part = nullptr;   // norm-backward sums of THIS layer's norm, per tile of the dgrad launch of the layer above that
  int nr_T = 0;               // produced them in its epilogue (ConvGemmParams::nr_*); nr_T = tiles per image, 0 = not folded
  int CoutK;                  // Cout padded to a multiple of 64 (channels of the dY frame)
  int CoutN;                  // rows per weight slab in the forward GEMM
  int CinN;                   // rows per weight slab in the dgrad GEMM
  FrameMap xmap, dymap;
  sg_bf16 *X, *Y, *dY;
  void* dX;  // bf16 except h0 (fp32)
  float* Yf32;
  int dxH, dxW, dx_oy, dx_ox, dx_fold, dx_f32;
  float *stats, *bsums;
  int* bsync;  // per-image arrival counters of the fused norm backward (in the zeroed-every-step range)
  float* stats_part;  // per-tile partial statistics written by the conv epilogue
  int stats_T, stats_T_max;
  sg_bf16 *Wf, *Wd;
  PackParams packf, packd;
  float* wscratch;
  int wscratch_elems;
  int unpack_mode;
  std::vector<ConvGemmLaunch> fwd, dgrad;
  std::vector<WgradLaunch> wgrad;
};

struct Net {
  std::vector<Layer> L;
  std::vector<TensorInfo> T;
  int64_t nparams;
  float *p, *g, *m, *v;
};

struct Arena {
  uint8_t* base;
  size_t off;
  void* take(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    void* r = base ? base + off : nullptr;
    off += bytes;
    return r;
  }
};

struct Engine {
  sggan_config cfg;
  cudaStream_t st;
  // weight-gradient GEMMs run on a side stream: they only need the layer's input frame and dY, so they fill the SMs
  // that the dgrad / norm-backward chain on `st` leaves idle at its wave tails; joined at the end of each phase
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_comm = nullptr;
  bool side_used = false;
  void join_side();
  bool dry;
  Net G, D;
  int Hd, Wd, Ho, Wo;  // D logit grid and broadcast output grid
  // misc buffers
  float *fake, *h4, *logits, *loss, *dD, *edge_w, *dGl;
  sg_bf16* resG[2];
  static constexpr size_t kMaxPackJobs = 64;
  PackParams* pack_jobs[2];
  int* pack_starts[2];
  int pack_njobs[2], pack_blocks[2];
  float* wg_part;  // split-K partial tiles of the weight-gradient GEMMs (shared by all layers)
  float* in_part;  // per-block partial sums of the norm-backward reduce pass (consumed by the apply pass that follows)
  float* red_scratch = nullptr;        // deposits of the ordered cross-block sums of the loss / seed kernels (OrderedSum)
  unsigned int* red_ticket = nullptr;  // one arrival counter per such kernel (self-resetting)
  int glue_err = 0;  // first failed row-stream launch of the current phase (reported when the phase ends)
  size_t wg_part_elems;
  uint8_t *zero_begin, *zero_end;
  int64_t step;
  long long* step_dev = nullptr;  // the same count on the device (Adam reads it: a captured graph must not bake t in)
  // captured sggan_train_step graphs (sggan_graph_capture), one per set of device pointers: double-buffered inputs
  // alternate between two of them
  struct StepGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
  };
  static constexpr int kMaxStepGraphs = 4;
  StepGraph graphs[kMaxStepGraphs];
  int ngraphs = 0, graph_sel = -1, graph_next = 0;
  int adam_mask = 0;  // which nets have been updated in the current step (bit 0 G, bit 1 D)
  int nlaunch;
  bool weights_ready;
  std::string err;
  // optional device-side timing of the dominant kernel (the 3x3 residual-block convolutions)
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  bool prof_on = false;
  int prof_kind = 0;  // 0 residual-block conv forward, 1 norm-apply forward of those layers, 2 their norm backward,
                      // 3 the generator-side loss kernels (one group per step)

  int build(const sggan_config& c, void* ws, size_t ws_bytes, cudaStream_t stream, bool dry_run, size_t* need);
  int pack_weights(int net, cudaStream_t s = nullptr);
  int upload_pack_jobs(int net);
  int gen_forward(const float* real_A, float* fake_out);
  int disc_forward_2b(const float* real_img, const float* fake_img, int nimg_each);
  int disc_forward_user(const float* x, const float* mask, float* logits_out);
  int step_fwd_bwd_d(const float* real_A, const float* seg_A, const float* mask, float* losses_out);
  int step_bwd_g(int part = -1);
  int bwd_split_layer() const;
  // state carried from part 0 to part 1 of a split generator backward
  int bwd_cur = 0;
  GradSrc bwd_gres = {}, bwd_add = {};
  bool bwd_half_done = false;
  int step_adam(int net, bool on_side_stream = false);

 private:
  int build_net_g();
  int build_net_d();
  void alloc_and_prepare(Net& n, Arena& a, bool zero_part);
  int prepare_layer(Net& n, int li);
  int run_conv_list(const std::vector<ConvGemmLaunch>& v);
  int run_wgrad(Layer& l, Net& n);
  void in_apply(Net& n, int li, sg_bf16* dst, const FrameMap& dmap, const sg_bf16* res, const FrameMap* rmap);
  void in_bwd(Net& n, int li, const GradSrc& g1, const GradSrc& g2, int nb_act, int act_wrap, int nb_param,
              sg_bf16* gather_dst = nullptr, int gH = 0, int gW = 0);
  GradSrc dx_src(const Layer& l) const;
  const float *real_A_, *seg_A_, *mask_;
  float* losses_out_;
};

}  // namespace sggan
