"""CPU oracle for the baseline-JPEG decode path -- TEST INFRASTRUCTURE ONLY.

`oracle.port`    : plain-C restatement of the reference algorithm (oracle/jpeg_oracle.c).
`oracle.refbind` : the real reference compiled from /root/reference (oracle/_ref/, built by
                   oracle/build_ref.sh where the reference is mounted).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (hls_jpeg_decoder_b200) never does.
"""
