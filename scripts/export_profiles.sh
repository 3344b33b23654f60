# Copies the evidence of scripts/gpu_final_profiles.sh from gpurun_out/ (scratch) into profiles/ (tracked):
# bench records, launch lists, and the text summaries of the ncu reports (details page as csv, per-source-line table)
# that gpu_final_profiles.sh produced on the GPU box.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out
P=profiles
TAG=${1:-r1_final}
cp $G/final_bench_der.json $P/${TAG}_bench_der.json
cp $G/final_bench_dif.json $P/${TAG}_bench_dif.json
cp $G/final_bench_reference_der.json $P/${TAG}_bench_reference_der.json
cp $G/final_bench_reference_dif.json $P/${TAG}_bench_reference_dif.json
cp $G/final_launches_staged_der.csv $P/${TAG}_launches_bench_der.csv
cp $G/final_launches_lm_der.csv $P/${TAG}_launches_lm_der_monolithic.csv
cp $G/final_launches_ekf.csv $P/${TAG}_launches_ekf_update.csv
for pair in staged_jac:staged_jac staged_solve:staged_solve staged_eval:staged_eval image8:image8_der \
            image8_dif:image8_dif lm_der:lm_der lm_dif:lm_dif ekf_cov:ekf_cov; do
  src=${pair%%:*}; dst=${pair##*:}
  [ -f $G/export/${src}_details.csv ] && cp $G/export/${src}_details.csv $P/${TAG}_ncu_full_${dst}_details.csv
  [ -f $G/export/${src}_by_line.txt ] && cp $G/export/${src}_by_line.txt $P/${TAG}_${dst}_by_line.txt
done
ls -la $P | grep $TAG
