for c in 12 16 20 24 32; do TDS_RASTER_CELL=$c python profiles/time_raster.py 2>&1 | tail -1 | sed "s/^/cell $c /"; done
