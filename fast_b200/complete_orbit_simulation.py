"""Orbit-sweep driver: host mirror of the reference's `fast/complete_orbit_simulation.py` (same
function names and arguments).  A satellite pass over a ground telescope is sampled in time and one
`Fast` simulation is configured per sample (range, zenith angle, azimuth, point-ahead angle and
downlink anisoplanatic angle); `fast_b200.sweep.run_sweep` then runs all of them with one host
synchronisation.

The pass geometry needs `skyfield` (third-party, not a dependency of the Monte-Carlo path): it is
imported lazily, and `FAST_sat_orbit(..., geometry=...)` accepts precomputed geometry instead, so
the driver itself works without it.  The field-of-view trigonometry is the pure function
`fov_offsets` below.
"""
import datetime

import numpy

from .fast import Fast


def _skyfield():
    try:
        from skyfield.api import load, wgs84
    except ImportError as e:                                     # pragma: no cover - optional dependency
        raise ImportError("skyfield is required for the orbit geometry (get_satellite_obj, get_sample_time, "
                          "get_angles_positions); pass geometry=... to FAST_sat_orbit to do without") from e
    return load, wgs84


def get_satellite_obj(TLE_file_path, satellite_name=None):
    '''skyfield satellite object from a TLE file (the named one, or the first)
    (fast/complete_orbit_simulation.py:9-27).'''
    load, _ = _skyfield()
    satellites = load.tle_file(TLE_file_path)
    if satellite_name != None:  # noqa: E711
        return {sat.name: sat for sat in satellites}[satellite_name]
    return satellites[0]


def get_sample_time(satellite, tele_lat, tele_lon, N=10, start=None, period=10, min_altitude_degrees=5.0,
                    max_altitude_degree=90.0, zenith_stop=False):
    '''N sample times [s] over the highest pass (culmination <= max_altitude_degree) within
    `period` days of `start`, from rise above min_altitude_degrees to set (or to culmination if
    zenith_stop), and the UTC datetime of the rise (fast/complete_orbit_simulation.py:29-95).'''
    load, wgs84 = _skyfield()
    ts = load.timescale()
    site = wgs84.latlon(tele_lat, tele_lon)
    t0 = ts.from_datetime(start) if start != None else satellite.epoch  # noqa: E711
    t1 = ts.from_datetime(t0.utc_datetime() + datetime.timedelta(days=period))
    times, events = satellite.find_events(site, t0, t1, min_altitude_degrees)
    best, best_alt = None, 0
    for i, ev in enumerate(events):
        alt = (satellite - site).at(times[i]).altaz()[0].degrees
        if ev == 1 and max_altitude_degree >= alt >= best_alt:
            best, best_alt = i, alt
    if best == None:  # noqa: E711
        raise Exception("The satellite doesn't pass over the telescop during the research period")
    i = best
    while i > 0 and events[i] != 0:
        i -= 1
    t_rise = times[i]
    if zenith_stop:
        t_fall = times[best]
    else:
        i = best
        while i < len(events) - 1 and events[i] != 2:
            i += 1
        t_fall = times[i]
    dt = (t_fall.utc_datetime() - t_rise.utc_datetime()).seconds
    return numpy.linspace(0, dt, N), t_rise.utc_datetime()


def fov_offsets(alt0, az0, alt1, az1):
    """(dx, dy) [deg] of direction (alt1, az1) as seen in the field of view of a telescope pointing
    at (alt0, az0), all angles in radians: great-circle separation alpha split along the vertical
    circle (dy) and across it (dx, signed by the azimuth difference)
    (fast/complete_orbit_simulation.py:150-166)."""
    z0, z1 = numpy.pi / 2 - alt0, numpy.pi / 2 - alt1
    cos_a = numpy.cos(z1) * numpy.cos(z0) + numpy.sin(z1) * numpy.sin(z0) * numpy.cos(az1 - az0)
    with numpy.errstate(invalid='ignore', divide='ignore'):
        sin_a = numpy.sqrt(1 - cos_a ** 2)
        cos_o = (numpy.cos(z1) - cos_a * numpy.cos(z0)) / (sin_a * numpy.sin(z0))
        sin_o = numpy.sqrt(1 - cos_o ** 2)
        sep = numpy.degrees(numpy.arccos(cos_a))
        return numpy.sign(numpy.degrees(az1) - numpy.degrees(az0)) * sin_o * sep, cos_o * sep


def get_angles_positions(sample_times, satellite, tele_lat, tele_lon, t_rise, Tloop, rotations=False):
    '''Point-ahead angle and downlink anisoplanatic angle [arcsec, (x, y) in the telescope field of
    view], altitude / azimuth [deg] and range [m] of the satellite at every sample
    (fast/complete_orbit_simulation.py:98-187).'''
    load, wgs84 = _skyfield()
    ts = load.timescale()
    celerity = 2.997925e8
    site = wgs84.latlon(tele_lat, tele_lon)
    n = len(sample_times)
    paa, dl = numpy.zeros((n, 2)), numpy.zeros((n, 2))
    altitudes, azimuts, distances, rot = (numpy.zeros(n) for _ in range(4))

    def at(obs, seconds):
        return (satellite - obs).at(ts.from_datetime(datetime.timedelta(seconds=seconds) + t_rise)).altaz()

    for i, t in enumerate(sample_times):
        alt0, az0, dist0 = at(site, t)
        altitudes[i], azimuts[i], distances[i] = alt0.degrees, az0.degrees, dist0.m
        # where the satellite will be when the uplink arrives, seen from where the telescope was
        two_way = 2 * dist0.m / celerity
        past_site = wgs84.latlon(tele_lat, tele_lon - 360 * two_way / (24 * 3600))
        alt_p, az_p, _ = at(past_site, t + two_way)
        alt_d, az_d, _ = at(site, t + Tloop)
        paa[i] = fov_offsets(alt0.radians, az0.radians, alt_p.radians, az_p.radians)
        dl[i] = fov_offsets(alt0.radians, az0.radians, alt_d.radians, az_d.radians)
        if rotations:
            z0, zd = numpy.pi / 2 - alt0.radians, numpy.pi / 2 - alt_d.radians
            cos_a = numpy.cos(zd) * numpy.cos(z0) + numpy.sin(zd) * numpy.sin(z0) * numpy.cos(az_d.radians - az0.radians)
            sin_a = numpy.sqrt(1 - cos_a ** 2)
            b0 = numpy.arccos((numpy.cos(zd) - numpy.cos(z0) * cos_a) / (sin_a * numpy.sin(z0)))
            b1 = numpy.arccos((numpy.cos(z0) - cos_a * numpy.cos(zd)) / (sin_a * numpy.sin(zd)))
            rot[i] = numpy.pi - b1 - b0
    paa = numpy.nan_to_num(paa * 3600, nan=0.0, posinf=numpy.inf, neginf=-numpy.inf)
    dl = numpy.nan_to_num(dl * 3600, nan=0.0, posinf=numpy.inf, neginf=-numpy.inf)
    if rotations:
        return paa, dl, altitudes, azimuts, distances, rot
    return paa, dl, altitudes, azimuts, distances


def FAST_sat_orbit(fast_params, simu_params, TLE_file, geometry=None):
    '''
    Sample a satellite pass over a telescope and configure one FAST simulation per sample
    (fast/complete_orbit_simulation.py:190-232).

    INPUTS :
        fast_params = [dict] - parameters of the FAST simulation (layers with CN2_TURB == 0 are dropped)
        simu_params = [dict] - satellite_name, telescop_lat, telescop_lon, N_sample, t0_research,
            research_window, altitude_min, altitude_max, zenith_stop
        TLE_file = [string] - path to a local or online TLE file
        geometry = optional (PAAs, aniso_dl, altitudes, azimuts, distances) as returned by
            get_angles_positions, to skip the skyfield computation

    OUTPUTS :
        dict 'simulation_<idx>' -> Fast object, plus 'altitudes'.  Run them with
        fast_b200.sweep.run_sweep([d[f'simulation_{i}'] for i in range(N)]).
    '''
    p = fast_params.copy()
    if geometry is None:
        satellite = get_satellite_obj(TLE_file, simu_params['satellite_name'])
        sample_times, t0 = get_sample_time(satellite, simu_params['telescop_lat'], simu_params['telescop_lon'],
                                           simu_params['N_sample'], simu_params['t0_research'],
                                           simu_params['research_window'], simu_params['altitude_min'],
                                           simu_params['altitude_max'], simu_params['zenith_stop'])
        geometry = get_angles_positions(sample_times, satellite, simu_params['telescop_lat'],
                                        simu_params['telescop_lon'], t0, p['TLOOP'])
    PAAs, aniso_dl, altitudes, azimuts, distances = (numpy.asarray(g) for g in geometry)
    keep = numpy.array(fast_params['CN2_TURB']) > 0
    for key in ('CN2_TURB', 'H_TURB', 'WIND_DIR', 'WIND_SPD'):
        p[key] = numpy.array(fast_params[key])[keep]
    out = {}
    for i, zenith in enumerate(90 - altitudes):
        p['L_SAT'] = distances[i]
        p['DTHETA'] = PAAs[i, :]
        p['ANISO_DL'] = aniso_dl[i, :]
        p['ZENITH_ANGLE'] = zenith
        p['AZIMUT_SAT'] = azimuts[i]
        out[f'simulation_{i}'] = Fast(dict(p))      # own copy: the sims must not share one params dict
    out['altitudes'] = altitudes
    return out


def FAST_sat(sat_apparent_speed, fast_params):
    '''One simulation whose downlink anisoplanatic angle is the apparent angular rate times the
    loop delay (fast/complete_orbit_simulation.py:234-236).'''
    fast_params['ANISO_DL'] = sat_apparent_speed * fast_params['TLOOP']
    return Fast(fast_params)
