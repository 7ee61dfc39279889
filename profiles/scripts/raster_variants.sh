# raster timing of experimental builds (torchdrivesim_b200/_build/libtds_<name>.so, see _build.build_library(out_path=...))
python profiles/time_raster.py
for v in "$@"; do TDS_B200_LIB=torchdrivesim_b200/_build/libtds_$v.so python profiles/time_raster.py; done
python profiles/time_raster.py
