python -m pytest tests/test_gpu_raster.py tests/test_gpu_rollout.py tests/test_gpu_graph.py -x -q 2>&1 | tail -3
for l in libtds_b200.so libtds_sr3.so libtds_sr7.so; do TDS_B200_LIB=$PWD/torchdrivesim_b200/_build/$l python profiles/time_raster.py; done
python profiles/time_raster.py
