"""`realtime_scripts` -- the frequency-domain ("FFT beamforming") backend of the reference's web
application (PC/application/realtime_scripts/), backed by libbf_b200.so.  Same module and function
names: beam_forming_algorithm.main(signal), calc_phase_shift_cartesian, calc_r_prime,
active_microphones, config."""
