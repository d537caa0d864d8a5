#!/bin/bash
# round-2 GPU visit 20: does the contact model survive a learner on the other tasks?  (ShadowHand: hand-object block-Jacobi; AnymalTerrain)
mkdir -p gpurun_out
timeout 500 python tools/train_ppo.py --task ShadowHand --num-envs 8192 --epochs 400 --horizon 8 --units 512,512,256,128 --lr 5e-4 --mini-epochs 5 --critic-coef 4 --out gpurun_out/r2_ppo_shadow_hand.json > gpurun_out/ppo_hand.log 2>&1; tail -1 gpurun_out/ppo_hand.log | cut -c1-330
grep -o '"epoch": [0-9]*, "env_steps": [0-9]*, "mean_episode_return": [^,]*, "mean_episode_length": [^,]*' gpurun_out/ppo_hand.log | awk 'NR%8==1' | tail -6
timeout 500 python tools/train_ppo.py --task AnymalTerrain --num-envs 4096 --epochs 300 --horizon 24 --minibatch 16384 --units 512,256,128 --lr 3e-4 --mini-epochs 5 --critic-coef 2 --out gpurun_out/r2_ppo_anymal_terrain.json > gpurun_out/ppo_anymal.log 2>&1; tail -1 gpurun_out/ppo_anymal.log | cut -c1-330
grep -o '"epoch": [0-9]*, "env_steps": [0-9]*, "mean_episode_return": [^,]*, "mean_episode_length": [^,]*' gpurun_out/ppo_anymal.log | awk 'NR%6==1' | tail -6
