// bh8_hud.h -- the reference's HUD text: cv::putText(img, text, {x, y}, cv::FONT_HERSHEY_PLAIN, 1, colour, 1)
// of blackhole_solution_test.cc:313-325 (camera position, basis vectors, field of view, drawn into the frame
// after rendering).  A blit over the glyph table of bh8_hud_font.h (rasterised from OpenCV itself): for text
// that lies inside the image the result equals cv::putText bit for bit (tests/test_hud.py).  Host code today
// (frames that come back to the host, bh8_draw_text); the sink's device-side blit uses the same table.
#ifndef BH8_HUD_H_
#define BH8_HUD_H_

#include <stddef.h>
#include <stdint.h>

#include "bh8_hud_font.h"

// Pixels outside the image are skipped (OpenCV clips the glyph's LINES instead, which can differ by a pixel
// for a glyph cut by the image border; the reference's HUD starts at x = 0 with every glyph inside).
static inline void bh8_hud_draw(uint8_t* bgr, int rows, int cols, size_t row_stride_bytes, int x, int y,
                                const char* text, uint8_t b, uint8_t g, uint8_t r) {
  for (const unsigned char* p = reinterpret_cast<const unsigned char*>(text); *p; ++p) {
    const int c = (*p < BH8_HUD_FIRST || *p > BH8_HUD_LAST) ? '?' : *p;  // putText draws '?' for these
    const uint16_t* glyph = bh8_hud_glyph[c - BH8_HUD_FIRST];
    for (int gr = 0; gr < BH8_HUD_CELL; ++gr) {
      const int yy = y + BH8_HUD_TOP + gr;
      const uint16_t bits = glyph[gr];
      if (!bits || yy < 0 || yy >= rows) continue;
      uint8_t* row = bgr + static_cast<size_t>(yy) * row_stride_bytes;
      for (int gx = 0; gx < BH8_HUD_CELL; ++gx) {
        const int xx = x + gx;
        if (!((bits >> gx) & 1) || xx < 0 || xx >= cols) continue;
        row[3 * xx] = b;
        row[3 * xx + 1] = g;
        row[3 * xx + 2] = r;
      }
    }
    x += bh8_hud_advance[c - BH8_HUD_FIRST];
  }
}

#endif  // BH8_HUD_H_
