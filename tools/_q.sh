python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python tools/profile_encode.py --bytes 268435456 --level 1 --iters 1 > gpurun_out/pe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_enc_match -c 1 -f -o gpurun_out/r2c_enc_match python tools/profile_encode.py --bytes 268435456 --level 1 --iters 1 > gpurun_out/ncu_e.log 2>&1
python tools/profile_decode.py --bytes 268435456 --corpus tick --chunk 4096 --iters 1 > gpurun_out/pd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_seq_t -c 3 -f -o gpurun_out/r2c_seq_4k python tools/profile_decode.py --bytes 268435456 --corpus tick --chunk 4096 --iters 1 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
