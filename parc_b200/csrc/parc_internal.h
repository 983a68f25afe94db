// Launch helpers shared between translation units of libparc_b200 (not part of the C ABI).
#pragma once

#include <stdint.h>

#include "../../include/parc_b200.h"

namespace parc {

// parc_body_loss with root positions read from rows `root_pos_stride` floats apart (body_loss.cu).
int body_loss_launch(const float* root_pos, int64_t root_pos_stride, const float* root_rot, const float* joint_rot,
                     const float* contacts, int64_t batch, int64_t frames, const ParcCharModel* model,
                     const ParcBodyPoints* pts, const ParcTerrainBatch* terrain, float w_pen, float w_contact,
                     float* pen_out, float* contact_out, float* g_root_pos, float* g_root_rot, float* g_joint_rot,
                     void* stream);

}  // namespace parc
