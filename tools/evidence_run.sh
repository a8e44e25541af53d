#!/bin/bash
# Produces the round's tracked evidence on a GPU box (run through gpurun from the repo root):
#   bench record, ncu launch list of two identical bench steps, `ncu --set full` captures summarised on the box
#   (the .ncu-rep files stay in /tmp: gpurun_out/ is limited to 64 MiB), a few A/B runs of tuning knobs.
T=${1:-r02i}
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 2 2 > gpurun_out/${T}_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_zgemm|k_zpass_g2r_tma|k_zpass_r2g|k_plane_vloc" --launch-skip 40 --launch-count 8 -o /tmp/${T}_hpsi -f python tools/profile_step.py 2 1 > gpurun_out/${T}_ncu_hpsi.log 2>&1
python tools/ncu_summary.py /tmp/${T}_hpsi.ncu-rep > gpurun_out/${T}_hpsi_full.md
python tools/ncu_summary.py --traffic 256 /tmp/${T}_hpsi.ncu-rep > gpurun_out/${T}_traffic_hpsi.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_zgemm|k_gjb_subpanel|k_gjb_prep" --launch-skip 60 --launch-count 6 -o /tmp/${T}_invert -f python tools/invert_one.py 8 > gpurun_out/${T}_ncu_invert.log 2>&1
python tools/ncu_summary.py /tmp/${T}_invert.ncu-rep > gpurun_out/${T}_invert_full.md
timeout 300 ncu --set full --clock-control none -k regex:"k_plane_rho_v2|k_shift_gemm|k_dots|k_seed_u|k_bicg_update|k_mr_mgs" --launch-skip 20 --launch-count 8 -o /tmp/${T}_misc -f python tools/profile_step.py 2 1 > gpurun_out/${T}_ncu_misc.log 2>&1
python tools/ncu_summary.py /tmp/${T}_misc.ncu-rep > gpurun_out/${T}_misc_full.md
# A/B of knobs (short benches, no Sigma_c / CPU legs)
for kv in SGW_ZG2R_DBT=1 SGW_PLANE_NT=256 SGW_PLANE_NT=320 SGW_PLANE_NT=352; do
  env $kv python bench.py --steps 4 --warmup 3 --no-sigma --no-cpu-baseline > gpurun_out/${T}_ab_${kv/=/_}.json 2>/dev/null
done
ls -la gpurun_out/
