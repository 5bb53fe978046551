"""Drop-in `groma` package surface backed by the B200-native engine (groma_b200).

Only the import paths the reference's callers use on the forward path are provided:
groma.model.groma.{GromaConfig, GromaModel}, groma.model.ddetr.{CustomDDETRConfig, CustomDDETRModel}, groma.constants
(SURVEY.md section 8b).  Training, data loading and serving are out of scope."""
