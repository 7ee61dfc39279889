for c in 4 5 6 8 10 12; do TDS_RASTER_CELL=$c python profiles/time_raster.py 2>&1 | tail -1 | sed "s/^/cell $c /"; done
