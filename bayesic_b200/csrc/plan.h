// The opaque plan object behind bb_plan_* (include/bayesic_b200.h).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/bayesic_b200.h"

struct bb_plan {
  std::vector<bb_node_desc> nodes;
  std::vector<int32_t> outputs;
  int32_t n_inputs = 0;
  int32_t last_launches = 0;
};
