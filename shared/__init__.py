"""Drop-in `shared` package: the reference's Python surface for the observation path
(disturbances_gpu, clip_ppo_utils, disturbance_types), backed by the sm_100a kernels."""
