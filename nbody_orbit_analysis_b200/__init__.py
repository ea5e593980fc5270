"""orbit-b200: B200-native orbit-tracking hot path of ``orbitanalysis``.

Sub-modules mirror the reference package layout so that
``import nbody_orbit_analysis_b200 as orbitanalysis`` drops in for the
tracking path: ``track_orbits``, ``track_orbits_onthefly``, ``progenitors``,
``postprocessing``, ``utils``.
"""
__version__ = '0.1.0'
