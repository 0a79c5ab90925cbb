"""Multi-output branching GP (mBGP): one learned branching point per output, one shared assignment q(Z)."""
