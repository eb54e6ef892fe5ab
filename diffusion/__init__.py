"""Import-path compatibility with the reference package layout (`diffusion.models.models.stable_diffusion_2` is the
`_target_` of yamls/hydra-yamls/SD-2-base-*.yaml, reference SD-2-base-256.yaml:14-15). Everything lives in diffusion_b200."""
