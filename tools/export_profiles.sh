# gpurun_out/<tag>_*.ncu-rep (tools/gpu_ncu.sh) -> tracked summaries profiles/<tag>_ncu_*.json.   usage: bash tools/export_profiles.sh r02c
TAG=$1
declare -A WHAT=(
 [s0_k11_c2]="generator stage 0 (C=256, T=6888/clip) ResBlock conv2, k=11, with residual"
 [s1_k3_c1]="generator stage 1 (C=128, T=55104/clip) ResBlock conv1, k=3"
 [s1_k11_c2]="generator stage 1 (C=128) ResBlock conv2, k=11, with residual"
 [s2_k3_c2]="generator stage 2 (C=64, T=110208/clip) ResBlock conv2, k=3, with residual"
 [s2_k3_pair]="generator stage 2 (C=64, T=110208/clip) fused ResBlock conv pair, k=3 (tcpair_kernel<64>)"
 [s3_k3_pair]="generator stage 3 (C=32, T=220416/clip) fused ResBlock conv pair, k=3 (tcpair_kernel<32>)"
 [s3_k7_c1]="generator stage 3 (C=32) ResBlock conv1, k=7"
 [s2_k11_c1]="generator stage 2 (C=64) ResBlock conv1, k=11"
 [s3_k3_c1]="generator stage 3 (C=32, T=220416/clip) ResBlock conv1, k=3"
 [s3_k3_c2]="generator stage 3 (C=32) ResBlock conv2, k=3, with residual"
 [s3_k11_c2]="generator stage 3 (C=32) ResBlock conv2, k=11, with residual"
 [wn_in]="posterior-encoder WaveNet in_layer (192 -> 384, k=5) with the gate epilogue"
)
for n in "${!WHAT[@]}"; do
  f=gpurun_out/${TAG}_$n.ncu-rep
  [ -f $f ] || continue
  ncu -i $f --page raw --csv 2>/dev/null > /tmp/raw_$$.csv
  python tools/summarize_ncu.py full /tmp/raw_$$.csv profiles/${TAG}_ncu_full_$n.json "ncu --set full --clock-control none, one launch, batch 8 x 10 s, f16x3: ${WHAT[$n]}" > /dev/null
  echo "profiles/${TAG}_ncu_full_$n.json"
done
[ -f gpurun_out/${TAG}_launches.csv ] && python tools/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv profiles/${TAG}_ncu_launches_b32_f16x3.json | head -8
[ -f gpurun_out/${TAG}_dram_b32.csv ] && python tools/summarize_ncu.py dram gpurun_out/${TAG}_dram_b32.csv profiles/${TAG}_ncu_dram_b32_f16x3.json | tail -8
rm -f /tmp/raw_$$.csv
