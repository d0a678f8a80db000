"""`oneka` -- the reference's package name, kept so that `from oneka.stochastic import
create_stochastic_capturezone` etc. keep working unchanged on top of onekapy_b200."""
