"""Patch files: the reference's persistence.Configuration JSON (persistence/persistence.go:6-88) read and written through
the signal.Chain / spatializer.Spatializer API, so that the Go host and this side configure their chains from the SAME file
(SURVEY.md 8d asks for CHAIN12 in this form: examples/chain12.patch.json).

persist() follows controller.persistenceSaveHandler (controller/controller.go:1195-1370), restore() follows the restore
branch of controller.persistenceRestoreHandler (controller/controller.go:944-1110), statement by statement; only the web
plumbing around them (multipart upload, HTTP response) is left out.  Works on anything with the Chain API of this package
(the product's gdg.Chain; the tests also drive the oracle's and the executed reference's chains with it).
"""
import json

UNIT_TYPES = ["signal_generator", "noise_gate", "bandpass", "auto_wah", "auto_yoy", "compressor", "octaver", "excess", "fuzz",
              "overdrive", "distortion", "tone_stack", "chorus", "flanger", "phaser", "tremolo", "ring_modulator", "delay",
              "reverb", "power_amp", "cabinet"]                 # effects.UnitTypes(): effects/effects.go:543-565
PARAMETER_TYPE_DISCRETE, PARAMETER_TYPE_NUMERIC = 1, 2            # effects/effects.go:12-16
BLOCK_SIZE = 8192                                                  # controller/controller.go:36


class PatchError(ValueError):
    pass


def persist(chains, spat=None, frames_per_period=BLOCK_SIZE, application="go-dsp-guitar", metronome=None):
    """controller.go:1195-1370 -> the persistence.Configuration as a dict (json.dumps gives the file).
    frames_per_period: BLOCK_SIZE in batch mode, hwio.FramesPerPeriod() when bound to JACK (controller.go:1200-1207).
    metronome: {"Master", "BeatsPerPeriod", "Speed", "TickSound", "TockSound"} or None (no metronome: zeros, :1335-1355)."""
    channels = []
    for chain_id, chain in enumerate(chains):
        units = []
        for unit_id in range(chain.length()):
            discrete, numeric = [], []
            for p in chain.parameters(unit_id):                                        # controller.go:1255-1289
                if p["Type"] == PARAMETER_TYPE_DISCRETE:
                    discrete.append({"Key": p["Name"], "Value": p["DiscreteValues"][p["DiscreteValueIndex"]]})
                elif p["Type"] == PARAMETER_TYPE_NUMERIC:
                    numeric.append({"Key": p["Name"], "Value": int(p["NumericValue"])})
            units.append({"Type": UNIT_TYPES[chain.unit_type(unit_id)], "Bypass": bool(chain.get_bypass(unit_id)),
                          "DiscreteParams": discrete, "NumericParams": numeric})
        if spat is not None:                                                           # controller.go:1305-1316
            sp = {"Azimuth": float(spat.get_azimuth(chain_id)), "Distance": float(spat.get_distance(chain_id)),
                  "Level": float(spat.get_level(chain_id))}
        else:
            sp = {"Azimuth": 0.0, "Distance": 0.0, "Level": 0.0}
        channels.append({"Units": units, "Spatializer": sp})
    m = metronome or {}
    return {"FileFormat": {"Application": application, "Type": "patch", "Version": {"Major": 1, "Minor": 0}},
            "FramesPerPeriod": int(frames_per_period), "Channels": channels,
            "Metronome": {"Master": bool(m.get("Master", False)), "BeatsPerPeriod": int(m.get("BeatsPerPeriod", 0)),
                          "Speed": int(m.get("Speed", 0)), "TickSound": str(m.get("TickSound", "")), "TockSound": str(m.get("TockSound", ""))}}


def restore(configuration, chains, spat=None):
    """controller.go:944-1110.  `configuration`: dict, JSON text or bytes.  Every chain named by the file is emptied and
    rebuilt: AppendUnit(type) -> SetDiscreteValue for every discrete parameter -> SetNumericValue for every numeric one ->
    SetBypass, in the file's order; then the channel's spatializer position.  Like the reference, a unit of unknown type is
    skipped and a setter's error does not stop the restore (the reference drops the setters' return values,
    controller.go:1078-1093); those errors are returned as a list of strings instead of being lost.
    Difference: a file with more channels than there are chains raises here - the reference reports it and then indexes past
    its chain slice (controller.go:1012-1031)."""
    if isinstance(configuration, (bytes, bytearray)):
        configuration = configuration.decode("utf-8")
    if isinstance(configuration, str):
        try:
            configuration = json.loads(configuration)
        except ValueError as e:
            raise PatchError("Error during unmarshalling: %s" % e)
    cfg = _fold(configuration)
    fmt = cfg.get("fileformat") or {}
    ver = fmt.get("version") or {}
    if fmt.get("type") != "patch":
        raise PatchError("Uploaded file is not a patch file.")                          # controller.go:973-981
    if int(ver.get("major", 0)) != 1 or int(ver.get("minor", 0)) < 0:
        raise PatchError("Incompatible version of file format.")                       # controller.go:983-991
    channels = cfg.get("channels") or []
    if len(channels) > len(chains):
        raise PatchError("WARNING: Restored file contains %d channels, but we currently have only %d. Restore may be incomplete."
                         % (len(channels), len(chains)))
    ignored = []

    def attempt(what, f, *a):
        try:
            f(*a)
        except Exception as e:                                                          # noqa: BLE001 - the reference ignores these
            ignored.append("%s: %s" % (what, e))
    for channel_id, channel in enumerate(channels):
        chain = chains[channel_id]
        while chain.length() > 0:                                                       # controller.go:1040-1047
            chain.remove_unit(chain.length() - 1)
        for unit in channel.get("units") or []:
            if unit.get("type") not in UNIT_TYPES:                                      # controller.go:1058-1071: not found -> skipped
                ignored.append("channel %d: unknown unit type %r" % (channel_id, unit.get("type")))
                continue
            chain.append_unit(UNIT_TYPES.index(unit["type"]))
            last = chain.length() - 1
            for p in unit.get("discreteparams") or []:
                attempt("channel %d unit %d %s" % (channel_id, last, p.get("key")), chain.set_discrete_value, last, p.get("key"), p.get("value"))
            for p in unit.get("numericparams") or []:
                attempt("channel %d unit %d %s" % (channel_id, last, p.get("key")), chain.set_numeric_value, last, p.get("key"), int(p.get("value", 0)))
            chain.set_bypass(last, bool(unit.get("bypass", False)))
        if spat is not None:                                                            # controller.go:1099-1106
            sp = channel.get("spatializer") or {}
            attempt("channel %d azimuth" % channel_id, spat.set_azimuth, channel_id, float(sp.get("azimuth", 0.0)))
            attempt("channel %d distance" % channel_id, spat.set_distance, channel_id, float(sp.get("distance", 0.0)))
            attempt("channel %d level" % channel_id, spat.set_level, channel_id, float(sp.get("level", 0.0)))
    return ignored


def _fold(v):
    """encoding/json matches object keys to struct fields case-insensitively: fold the keys, keep the values"""
    if isinstance(v, dict):
        return {str(k).lower(): _fold(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_fold(x) for x in v]
    return v
