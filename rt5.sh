for v in "" big2 big3; do
  if [ -n "$v" ]; then export TDS_B200_LIB=torchdrivesim_b200/_build/libtds_$v.so; fi
  echo "== $v"; python profiles/time_raster_res.py 128 256 128; python profiles/time_raster_res.py 256 64 128
done
