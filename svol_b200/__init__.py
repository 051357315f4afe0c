"""svol_b200 -- B200-native (sm_100a) implementation of SVOL's head forward + Hungarian matching +
set losses behind the reference's ``lib/modeling`` API.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
