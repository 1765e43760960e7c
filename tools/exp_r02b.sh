# Gram chunk-depth A/B, ncu captures of the Gram kernel (bench size) and of the two DF kernels (config 3)
set -x
for lay in 0 1; do
  python tools/time_gram.py 393216 10000 $lay
  PODSDF_LIB=build/variants/libpodsdf_gkb32.so python tools/time_gram.py 393216 10000 $lay
done
python tools/time_gram.py 393216 10000 0 && ncu --set full --clock-control none -k regex:gram_dmma -s 1 -c 1 -o /tmp/gram python tools/time_gram.py 393216 10000 0 > gpurun_out/ncu_gram.log 2>&1
python tools/time_chunks.py 1256 && ncu --set full --clock-control none -k regex:"df_time|df_plane" -s 4 -c 2 -o /tmp/df python tools/time_chunks.py 1256 > gpurun_out/ncu_df.log 2>&1
for n in gram df; do
  ncu -i /tmp/$n.ncu-rep --page raw --csv > gpurun_out/r02_${n}_raw.csv
  ncu -i /tmp/$n.ncu-rep --page details > gpurun_out/r02_${n}_details.txt
done
ls -la gpurun_out/ | tail -8
