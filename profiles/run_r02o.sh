SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_phases3.so timeout 300 python profiles/phase_profile3.py 2>&1 | tail -16
