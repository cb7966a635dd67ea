"""CPU oracle of the stereo point+line front-end hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import, link or
execute it, and only as the checker (or as the timed CPU baseline), never as the path measured or
shipped.  The product path is pl-slam_b200/ (CUDA) and fails loudly without a GPU.

Parity status: the reference's own implementation of this path cannot be
built as a whole here (stvo-pl is not vendored; no OpenCV/Eigen C++ packages).  The oracle is
  * OpenCV 4.13 (python cv2, same arithmetic library the reference links) for ORB / LSD / kNN,
  * C and numpy restatements of the vendored LBD, the in-tree Gauss-Newton and the stvo-pl
    matcher/stereo logic, each citing the reference file:line it follows,
  * and, to check the LBD / KeyLine restatement, the vendored 3rdparty/line_descriptor sources
    themselves, compiled unmodified into oracle/_ref/ (oracle/ref_build/, oracle/refbin.py).
The reference holds no golden vectors or tests for this path => for the stvo-pl
pieces, the Gauss-Newton twin and the map-feature helper this is "parity unpinned"; the cv2-backed
pieces are pinned against cv2 itself, the LBD / KeyLine stage against the reference's own code.
"""
