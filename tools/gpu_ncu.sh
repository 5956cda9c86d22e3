# ncu captures of the tensor-core conv kernels as they run today.  usage: bash tools/gpu_ncu.sh <tag> [wide_variant]
TAG=${1:-r02a}
export OVC_WIDE_VARIANT=${2:-3}
mkdir -p gpurun_out
CAPS=${CAPS:-"s0_k11_c2:111 s1_k3_c1:117 s1_k11_c2:130 s2_k3_pair:136 s2_k11_c1:145 s3_k3_pair:152 s3_k7_c1:155 s3_k11_c2:162 wn_in:40"}
if [ -z "$ONLY_CAPS" ]; then
# every launch of one call (second call), device time per launch
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 208 -c 208 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python tools/ncu_target.py --batch 32 --calls 2 > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
# DRAM bytes of every tensor-core conv launch of the second call at the BENCHED batch (single pass: no replay, no save/restore)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:tcconv|tcpair' -s 167 -c 167 \
  --csv --log-file gpurun_out/${TAG}_dram_b32.csv python tools/ncu_target.py --batch 32 --calls 2 > gpurun_out/${TAG}_dram_b32.log 2>&1; echo "dram pass rc=$?"
fi
# --set full on chosen conv launches of the second call (tcconv launch index within a call, see tools/ncu_target.py)
cap() {  # name index
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tcconv|tcpair' -s $((167 + $2)) -c 1 -f -o gpurun_out/${TAG}_$1 \
    python tools/ncu_target.py --batch 8 --calls 2 > gpurun_out/${TAG}_$1.log 2>&1; echo "ncu $1 rc=$?"
}
for c in $CAPS; do cap ${c%%:*} ${c##*:}; done
ls -la gpurun_out | head -40
