// Tag API entries: "experimental" (docstring note), "deprecated", and "sm_100a" for entries backed by hand-written kernels.
document.addEventListener("DOMContentLoaded", function () {
  function badge(dt, cls, text) {
    var s = document.createElement("span");
    s.className = "sdl-badge " + cls;
    s.textContent = text;
    dt.appendChild(s);
  }
  document.querySelectorAll("dl.py > dt").forEach(function (dt) {
    var dd = dt.nextElementSibling;
    if (!dd) { return; }
    var text = dd.textContent.toLowerCase();
    if (text.indexOf("this api is experimental") >= 0) { badge(dt, "experimental", "Experimental"); }
    if (text.indexOf("deprecated") >= 0) { badge(dt, "deprecated", "Deprecated"); }
    if (text.indexOf("sm_100a") >= 0) { badge(dt, "b200", "sm_100a kernel"); }
  });
});
