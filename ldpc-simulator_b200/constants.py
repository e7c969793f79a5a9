"""Seeds of the reference's LCG noise sources (python_ldpc_app/constants.py:2-3).
Only channel modes 2/3 used them; kept so that ``Channel.gen_ptr`` objects look the same."""
IDUM1 = 83685
IDUM2 = 11111
