// Compact observation record for the host-buffer path (mdr_step_host with an MdrHostCtx): with the default
// observation layout (implicit `neighbours` messages, utils.py:842-868), 4*C of the 11 + 4*C floats of a row are copies
// of OTHER houses' 4-float messages.  Instead of sending E*N*F reals over PCIe, a house sends 16:
//   [0..10] its own features, [11..14] its message (dT/5, sso, P/7500, Pmax/7500), [15] 1 / lockout_duration
// (the receiver scales the sender's sso by its OWN lockout duration, utils.py:848-850), and the rows are expanded on
// the host (mdr_host.cu).  Same arithmetic, operation for operation, as the step kernels' row assembly, on the state
// the step just stored: the expanded rows are bit-identical to the rows a step kernel writes.
// Included by mdr_kernels.cu inside namespace mdr; not a standalone translation unit.
#pragma once

template <typename R>
__global__ void __launch_bounds__(256) compact_obs_kernel(const __grid_constant__ KernelParams p, R* __restrict__ out) {
  using T2 = typename Vec<R>::T2;
  using T4 = typename Vec<R>::T4;
  const unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= (unsigned)p.E * (unsigned)p.N) return;
  const int e = (int)(h / (unsigned)p.N);
  const T2 tt = reinterpret_cast<const T2*>(p.temps)[h];
  const T4 cb = reinterpret_cast<const T4*>(p.coef_b)[h];
  const T2 cc = reinterpret_cast<const T2*>(p.coef_c)[h];
  const int hv = p.hvac[h];
  const R target = cb.w, p_on = cb.z;
  const int on = hv & 1, lock = (hv >> 1) & 1, sso = hv >> 2;
  const R inv_lock = inv_real(cc.y);
  const R inv_norm = (R)p.inv_norm_reg_sig;
  const R pw = on ? p_on : (R)0;
  T4* o = reinterpret_cast<T4*>(out) + (size_t)h * 4;
  o[0] = make4((tt.x - 20) * (R)0.2, (tt.y - 20) * (R)0.2, (target - 20) * (R)0.2, cc.x);
  o[1] = make4(p_on * (R)p.cop_over_def_cap, (R)on, (R)lock, (R)sso * inv_lock);
  o[2] = make4((R)1, (R)(p.signal[e] * p.inv_norm_sig_agents), (R)(p.cluster_power[e] * p.inv_norm_sig_agents),
               (tt.x - target) * (R)0.2);
  o[3] = make4((R)sso, pw * inv_norm, p_on * inv_norm, inv_lock);
}

cudaError_t launch_compact_obs(const KernelParams& kp, int precision, void* out, cudaStream_t stream) {
  const size_t total = (size_t)kp.E * kp.N;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (precision == MDR_F32) compact_obs_kernel<float><<<blocks, 256, 0, stream>>>(kp, reinterpret_cast<float*>(out));
  else compact_obs_kernel<double><<<blocks, 256, 0, stream>>>(kp, reinterpret_cast<double*>(out));
  return cudaGetLastError();
}
