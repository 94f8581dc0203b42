// The other custom identification networks of the reference's ModelFetcher (V100, V110, V119, V200;
// T/python/visual_identification_network_torch.py:30-181,262-386,537-567) as one layer-list executor on fp32 CUDA cores.
#pragma once
#include "common.h"

#include <map>
#include <string>
#include <vector>

namespace tb {

struct ViNet;

// arch: tb_vi_config.arch (1 v100, 2 v110, 3 v119, 4 v200)
int vinet_create(ViNet **out, int arch, int channels, int num_classes, int max_images, std::vector<void *> &allocs);
void vinet_destroy(ViNet *n);
int vinet_commit(ViNet *n, const std::map<std::string, std::vector<float>> &sd);
int vinet_forward(ViNet *n, const uint8_t *img, int n_max, const uint32_t *n_dev, float *probs, float *logits,
                  uint32_t *top_id, float *top_p, cudaStream_t s, EventRing<5> &prof, uint64_t &launches);

}  // namespace tb
