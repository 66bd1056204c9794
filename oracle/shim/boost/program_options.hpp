// Minimal stand-in so the reference's density translation units compile without Boost.
// The hot-path TUs only need the *declaration* of variables_map (Density::main signature).
#pragma once
namespace boost { namespace program_options { class variables_map; } }
