# ncu captures of the K-time kernel (pair kernel) with and without the Philox rounds; summaries only (reps are large)
set -x
PODSDF_TIME_VARIANT=0 python tools/time_chunks.py 1256 && PODSDF_TIME_VARIANT=0 ncu --set full --clock-control none -k regex:df_time -s 2 -c 1 -o /tmp/v0 python tools/time_chunks.py 1256 > gpurun_out/ncu_v0.log 2>&1
PODSDF_LIB=build/variants/libpodsdf_r0.so PODSDF_TIME_VARIANT=0 python tools/time_chunks.py 1256 && PODSDF_LIB=build/variants/libpodsdf_r0.so PODSDF_TIME_VARIANT=0 ncu --set full --clock-control none -k regex:df_time -s 2 -c 1 -o /tmp/r0 python tools/time_chunks.py 1256 > gpurun_out/ncu_r0.log 2>&1
ls -la /tmp/*.ncu-rep
for n in v0 r0; do
  ncu -i /tmp/$n.ncu-rep --page raw --csv > gpurun_out/r02_ktime_${n}_raw.csv
  ncu -i /tmp/$n.ncu-rep --page details > gpurun_out/r02_ktime_${n}_details.txt
  ncu -i /tmp/$n.ncu-rep --page source --csv > gpurun_out/r02_ktime_${n}_source.csv
done
ls -la gpurun_out/
