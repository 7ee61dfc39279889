python -m pytest tests/test_gpu_raster.py -x -q 2>&1 | tail -2
python profiles/time_raster.py
for v in minb5 minb8; do TDS_B200_LIB=torchdrivesim_b200/_build/libtds_$v.so python profiles/time_raster.py; done
