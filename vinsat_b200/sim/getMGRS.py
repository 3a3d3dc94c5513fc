"""MGRS grid-zone lookup used by SatCam.get_region: zone id -> (west, south, east, north) in degrees.

Same table as the reference's ``sim/getMGRS.py:5-30`` (and the same key ORDER, band by band from 80 S, zones west to
east inside a band: `get_region` returns the first zone whose closed box contains a point, so the order decides ties
on zone borders).  Built from the definition of the grid rather than by patching a regular grid afterwards.
"""

# 8-degree latitude bands from 80 S upwards (I and O are not used); the last band, X, is 12 degrees tall (72..84 N)
_BANDS = "CDEFGHJKLMNPQRSTUVWX"
# Norway (band V) and Svalbard (band X) exceptions: zone -> (west, east); a missing zone is None
_IRREGULAR = {
    "V": {31: (0, 3), 32: (3, 12)},
    "X": {31: (0, 9), 32: None, 33: (9, 21), 34: None, 35: (21, 33), 36: None, 37: (33, 42)},
}


def getMGRS():
    grid = {}
    for k, band in enumerate(_BANDS):
        south = 8 * k - 80
        north = 84 if band == "X" else south + 8
        special = _IRREGULAR.get(band, {})
        for zone in range(1, 61):
            west_east = special.get(zone, (6 * zone - 186, 6 * zone - 180))
            if west_east is None:
                continue
            grid["%02d%s" % (zone, band)] = (west_east[0], south, west_east[1], north)
    return grid
