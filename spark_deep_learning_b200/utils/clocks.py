"""Sample ``nvidia-smi`` clocks / throttle reasons while a benchmark's timed region runs
(B200_PROFILING.md "clocks DURING the timed region")."""
import shutil
import statistics
import subprocess

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler(object):
    def __init__(self, gpu_index=0, period_ms=200):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self._p = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return self
        try:
            self._p = subprocess.Popen(
                [exe, "--query-gpu=" + _QUERY, "--format=csv,noheader,nounits", "-lms",
                 str(self.period_ms), "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self._p = None
        return self

    def stop(self):
        """Returns ``{"sm_mhz": median, "sm_max_mhz": .., "reasons": [...], "samples": n}``."""
        if self._p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._p.terminate()
        try:
            out, _ = self._p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self._p.kill()
            out, _ = self._p.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # "under load": the upper half of the samples (idle samples before/after the loop drop out)
        loaded = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax),
                "power_w_max": max(power), "reasons": sorted(reasons), "samples": len(sm)}
