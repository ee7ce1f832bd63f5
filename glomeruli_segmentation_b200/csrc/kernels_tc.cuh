// tcgen05 / TMEM (fp16 operands, fp32 accumulate) kernels of the ESP blocks -- placeholder until the
// tensor-core path lands; espnet_set_mode(ESPNET_MODE_F16TC) reports it as unavailable.
#pragma once
namespace espnet {
inline bool tc_path_available() { return false; }
}  // namespace espnet
