// Device-resident rollout collection: the per-step action draw of the reference's learners,
//   c = Categorical(action_prob); action = c.sample(); return action.item(), action_prob[:, action.item()].item()
// (agents/ppo.py:68-75, once per agent and step, on the host), as ONE kernel over the whole [M, A] batch of action
// probabilities: inverse-CDF draw from a Philox uniform keyed by (row, draw index), the uint8 action written where the
// step kernel reads it and the chosen probability written into the rollout storage (`a_log_prob` of agents/ppo.py:92-107).
// Included by mdr_kernels.cu inside namespace mdr; not a standalone translation unit.
#pragma once

enum : uint32_t { STREAM_SAMPLE = 6 };

template <int kA>  // kA > 0: number of actions fixed at compile time (2 = on/off)
__global__ void __launch_bounds__(256) sample_actions_kernel(const float* __restrict__ probs, long long n_rows, int n_actions,
                                                             uint64_t seed, uint64_t draw_index,
                                                             const uint64_t* __restrict__ draw_counter,
                                                             uint8_t* __restrict__ actions, float* __restrict__ chosen_prob) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int A = kA > 0 ? kA : n_actions;
  const uint64_t draw = draw_index + (draw_counter != nullptr ? *draw_counter : 0);
  const uint4 r = philox4x32((uint32_t)row, (uint32_t)(row >> 32) ^ (uint32_t)(draw >> 32), (uint32_t)draw, STREAM_SAMPLE, seed);
  const float u = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
  const float* pr = probs + row * A;
  int a = 0;
  float pa;
  if (kA == 2) {
    const float2 p2 = *reinterpret_cast<const float2*>(pr);
    // Categorical normalises by the row sum (torch.distributions.Categorical(probs))
    a = u * (p2.x + p2.y) < p2.x ? 0 : 1;
    pa = a ? p2.y : p2.x;
  } else {
    float total = 0.0f;
    for (int i = 0; i < A; ++i) total += pr[i];
    const float target = u * total;
    float cum = 0.0f;
    a = A - 1;
    for (int i = 0; i < A; ++i) {
      cum += pr[i];
      if (target < cum) { a = i; break; }
    }
    pa = pr[a];
  }
  actions[row] = (uint8_t)a;
  if (chosen_prob != nullptr) chosen_prob[row] = pa;
}

cudaError_t launch_sample_actions(const float* probs, long long n_rows, int n_actions, uint64_t seed, uint64_t draw_index,
                                  const uint64_t* draw_counter, uint8_t* actions, float* chosen_prob, cudaStream_t stream) {
  const int threads = 256;
  const unsigned blocks = (unsigned)((n_rows + threads - 1) / threads);
  if (n_actions == 2 && (reinterpret_cast<uintptr_t>(probs) & 7) == 0)
    sample_actions_kernel<2><<<blocks, threads, 0, stream>>>(probs, n_rows, n_actions, seed, draw_index, draw_counter, actions, chosen_prob);
  else
    sample_actions_kernel<0><<<blocks, threads, 0, stream>>>(probs, n_rows, n_actions, seed, draw_index, draw_counter, actions, chosen_prob);
  return cudaGetLastError();
}
