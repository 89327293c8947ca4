"""``gto`` package surface of IRVLUTD/GraspTrajOpt on the B200 solver (host side, NumPy + ctypes)."""
